/* gcnb200.h -- C ABI of libgcnb200.so: the B200 (sm_100a) kernels behind geographconv's GCN
 * hot path.  This is the drop-in boundary: the reference has no FFI of its own
 * (it is pure Python on Theano/Lasagne), so the entry points below are the operations its
 * gcnmodel.py asks Theano for, one function per op, each citing the reference call site it
 * replaces.  geographconv_b200/gcnmodel.py binds them with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / numpy types.
 *  - every `const float*` / `float*` matrix is DEVICE memory, row-major fp32, with a leading
 *    dimension (`ld*`, in floats) that is a multiple of 4 and >= the logical width rounded up
 *    to 4; base pointers are 16-byte aligned.  Padding columns hold zeros and kernels keep them
 *    zero.  Index arrays are int32 (gcnmain.py:167-168, gcnmodel.py:329).
 *  - the CALLER owns all device memory (allocated with any allocator; the Python host uses
 *    torch as the HBM allocator).  The library never allocates or frees device memory; scratch
 *    is a caller-provided workspace (gcnb_set_workspace).
 *  - all ops are asynchronous on the context's stream; only gcnb_sync / gcnb_prof_collect /
 *    the *_sync helpers block the host.
 *  - return value: 0 (GCNB_OK) or a negative gcnb_status; gcnb_last_error(ctx) explains.
 *  - one host thread per context (the reference is single-threaded, gcnmodel.py:429-430).
 */
#ifndef GCNB200_H
#define GCNB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCNB_VERSION 100

typedef struct gcnb_ctx gcnb_ctx;

enum gcnb_status {
  GCNB_OK = 0,
  GCNB_E_INVALID = -1,     /* bad argument (shape, alignment, null pointer) */
  GCNB_E_CUDA = -2,        /* a CUDA runtime / driver call failed */
  GCNB_E_WORKSPACE = -3,   /* workspace missing or too small */
  GCNB_E_UNSUPPORTED = -4  /* shape outside what the kernels are instantiated for */
};

/* lasagne.nonlinearities used by gcnmodel.py: tanh (:347, live), rectify (:345, commented
 * out), sigmoid (:286 gate), linear (:188), selu (:290 residual_dense; element-wise epilogues of the SpMM
 * and the CUDA-core GEMM only -- the tcgen05 GEMM declines it). */
enum gcnb_act { GCNB_ACT_LINEAR = 0, GCNB_ACT_TANH = 1, GCNB_ACT_RELU = 2, GCNB_ACT_SIGMOID = 3, GCNB_ACT_SELU = 4 };

/* profiling classes: the step-time split SURVEY.md 8d asks for */
enum gcnb_tag {
  GCNB_TAG_SPMM_A = 0,  /* A_hat . H     (structured_dot(A, .), gcnmodel.py:130,153) */
  GCNB_TAG_SPMM_X = 1,  /* X . W0        (structured_dot(input, W), gcnmodel.py:39) */
  GCNB_TAG_SPMM_XT = 2, /* X^T . dz      (grad of gcnmodel.py:39 wrt W) */
  GCNB_TAG_GEMM = 3,    /* T.dot         (gcnmodel.py:126,149,285) and its grads */
  GCNB_TAG_ELEM = 4,    /* bias / activation / gate / dropout element-wise work */
  GCNB_TAG_LOSS = 5,    /* softmax cross-entropy, argmax, gathers (gcnmodel.py:376-389) */
  GCNB_TAG_ADAM = 6,    /* lasagne.updates.adam (gcnmodel.py:407) */
  GCNB_TAG_COPY = 7,    /* host<->device copies issued through this ABI */
  GCNB_TAG_SPMM_A_NARROW = 8, /* A_hat . (x Wout) and its gradient: K = classes, not the hidden width */
  GCNB_TAG_COMM = 9,    /* multi-GPU exchange kernels of this library: slice pushes (rows -> column slices) */
  GCNB_TAG_SYNC = 10,   /* peer barriers: launch + flag round trip + waiting for the slowest rank */
  GCNB_NTAGS = 11
};

/* ---------------------------------------------------------------- context ------------- */
int gcnb_version(void);
/* `stream` is a cudaStream_t (NULL = create a private non-blocking stream). */
int gcnb_create(int device, void* stream, gcnb_ctx** out);
int gcnb_destroy(gcnb_ctx* ctx);
const char* gcnb_last_error(const gcnb_ctx* ctx);
int gcnb_set_stream(gcnb_ctx* ctx, void* stream);
void* gcnb_get_stream(const gcnb_ctx* ctx);
int gcnb_set_workspace(gcnb_ctx* ctx, void* dev_ptr, size_t bytes);
/* named integer options: "spmm_variant" (0 = LDG.128 register gather, 1 = bulk-copy/TMA staged, 2 = L2-resident
 * column panels), "spmm_panel" (panel width of engine 2 in floats: 16, 32, 64), "spmm_panel_policy" (panel gathers: 0 =
 * default cache policy, 1 = L2 evict_last hint (default), 2 = evict_last with L1 allocation),
 * "spmm_unroll" (nonzeros gathered per batch), "gemm_tc" (1 = tcgen05 path where supported, the
 * default), "tc_launches" (read-only count of tcgen05 kernels launched), "sm_margin" (SMs the persistent
 * SpMM kernel leaves to concurrently running collectives), "spmm_sliced_engine" (gather engine of gcnb_spmm_csr_sliced_f32: -1 = by
 * operand size, the default; 0 / 1 / 2 as above), "prof_mask" (bit t set: ops of gcnb_tag t are timed while
 * profiling is on; default all), "peer_timeout_s" (gcnb_peer_barrier). */
int gcnb_set_option(gcnb_ctx* ctx, const char* name, int value);
int gcnb_get_option(const gcnb_ctx* ctx, const char* name, int* value);
int gcnb_sync(gcnb_ctx* ctx);
int gcnb_sm_count(const gcnb_ctx* ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
long long gcnb_launch_count(const gcnb_ctx* ctx);
/* per-tag CUDA-event timing of every op, recorded on the context's stream */
int gcnb_prof_enable(gcnb_ctx* ctx, int on);
int gcnb_prof_reset(gcnb_ctx* ctx);
/* synchronises the stream; fills ms[GCNB_NTAGS] and ops[GCNB_NTAGS] (accumulated since reset) */
int gcnb_prof_collect(gcnb_ctx* ctx, float* ms, long long* ops);

/* ---------------------------------------------------------------- transfers ----------- */
int gcnb_h2d(gcnb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int gcnb_d2h(gcnb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int gcnb_memset(gcnb_ctx* ctx, void* dst_dev, int byte, size_t bytes);
/* device-to-device copy of a rows x cols fp32 block between matrices with different leading dimensions (a column
 * panel of an activation made contiguous for the all-gather of the row-partitioned exchange) */
int gcnb_copy2d_f32(gcnb_ctx* ctx, const float* src, int32_t ld_src, float* dst, int32_t ld_dst, int32_t rows,
                    int32_t cols);

/* dst[i] = src[i] widened to int32 (device arrays, 16-byte aligned).  Column ids of a matrix with at most 65536 columns
 * (the BoW features: 10k-50k terms) cross PCIe as uint16 -- a quarter of the CSR bytes less -- and are widened here. */
int gcnb_expand_u16_i32(gcnb_ctx* ctx, const uint16_t* src, int64_t n, int32_t* dst);

/* ---------------------------------------------------------------- CSR ----------------- */
/* A CSR matrix resident in device memory plus its work decomposition ("plan"): every row is
 * cut into items of at most `chunk` nonzeros; a row that needs more than one item is a "long"
 * row whose items write partial sums that a fix-up pass adds in order (deterministic).
 * Replaces the scipy CSR that theano.sparse.csr_matrix inputs carry (gcnmodel.py:338,342). */
typedef struct gcnb_csr {
  int32_t n_rows;
  int32_t n_cols;
  int64_t nnz;
  const int32_t* rowptr; /* device, n_rows + 1 */
  const int32_t* colidx; /* device, nnz */
  const float* val;      /* device, nnz */
  const int32_t* items;  /* device, 4 * n_items: {row, begin, end, slot (-1 = direct)} */
  int32_t n_items;
  const int32_t* long_rows; /* device, 3 * n_long: {row, first_slot, n_slots_of_row} */
  int32_t n_long;
  int32_t n_slots;
  int32_t tag; /* gcnb_tag the SpMM time is booked under */
  /* gather engine for this matrix: 0 = LDG.128 register gather (best when the whole dense operand is L2
   * resident), 1 = bulk-copy (TMA engine) staged gather with persistent CTAs (best when the gathered rows
   * must come from HBM), 2 = L2-resident column panels (the product is computed 32 columns of B at a time,
   * panel-major, so B leaves HBM once per product; best when one n_cols x 128 B panel fits L2, e.g. A_hat.H
   * at N = 500k), -1 = the context's "spmm_variant" option, -2 = choose per call from the operand size and K */
  int32_t engine;
  int32_t unroll; /* nonzeros gathered per batch by engine 0 (2, 4, 8); 0 = the context's / auto */
} gcnb_csr;

/* Host-side planning over a HOST rowptr.  Call once with items == NULL to get the counts, then
 * again with host buffers of 4*n_items and 3*n_long int32 to fill. */
int gcnb_csr_plan(const int32_t* host_rowptr, int32_t n_rows, int32_t chunk, int32_t* n_items,
                  int32_t* n_long, int32_t* n_slots, int32_t* items, int32_t* long_rows);

/* fused epilogue of the SpMM: out = dropout(act(acc + bias)) or softmax(acc + bias);
 * C = out, or C += out when `accumulate`. */
typedef struct gcnb_epilogue {
  const float* bias;   /* device, K floats padded to a multiple of 4 with zeros; or NULL */
  int32_t act;         /* gcnb_act */
  int32_t softmax;     /* 1: row softmax over the K columns after the bias (fused up to K = 512; wider rows, e.g. the 930
                        * classes of TwitterWorld, take a row pass of their own after the product) */
  int32_t accumulate;  /* 0: C = out.  1: C += out.  2: C = epilogue(C + A.B): the product is added to what C
                        * already holds BEFORE bias / activation / dropout / softmax */
  float dropout_p;     /* 0: none.  Inverted dropout, scale 1/(1-p) (lasagne DropoutLayer) */
  uint64_t seed;       /* Philox key */
  int64_t row0;        /* global index of local row 0 (row-partitioned runs draw the same mask) */
  float* logits;       /* optional (softmax only): pre-softmax values, same ld as C; or NULL */
} gcnb_epilogue;

/* C[n_rows x K] = epilogue( A[n_rows x n_cols] . B[n_cols x K] )
 * theano.sparse.structured_dot: gcnmodel.py:39 (X.W0), :130 (A.(xW)), :153 (output layer) and
 * its gradient structured_dot(a.T, g).  `epi` may be NULL (plain product). */
int gcnb_spmm_csr_f32(gcnb_ctx* ctx, const gcnb_csr* A, const float* B, int32_t ldb, float* C,
                      int32_t ldc, int32_t K, const gcnb_epilogue* epi);
/* the gather engine (0, 1, 2) the call above would use for this matrix, leading dimension of B and K */
int gcnb_spmm_engine_for(const gcnb_ctx* ctx, const gcnb_csr* A, int32_t ldb, int32_t K);
/* workspace bytes the call above needs for this matrix and K */
size_t gcnb_spmm_workspace_bytes(const gcnb_csr* A, int32_t K);

/* ---------------------------------------------------------------- dense --------------- */
/* C[M x N] = act(op(A) . op(B) + bias), or C += op(A).op(B) when accumulate (bias/act then
 * unused).  op(A) is M x K: transA == 0 -> A is M x K row-major (lda), else A is K x M.
 * T.dot: gcnmodel.py:126,149,215 and the dgrad / wgrad products of its gradient. */
int gcnb_gemm_f32(gcnb_ctx* ctx, int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K,
                  const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
                  int32_t accumulate, const float* bias, int32_t act);
size_t gcnb_gemm_workspace_bytes(int32_t transA, int32_t M, int32_t N, int32_t K);
/* C[M x N] (+)= A1 . op(B1) + A2 . op(B2), both products M x K by K x N, in one pass over C (two k-loops into one
 * accumulator): the two dgrad products of a highway layer, dx += dTpre.Wt^T + V.Wh^T (gradient of gcnmodel.py:126,285).
 * Workspace: 2 x gcnb_gemm_workspace_bytes(0, M, N, K). */
int gcnb_gemm_pair_f32(gcnb_ctx* ctx, int32_t transB, int32_t M, int32_t N, int32_t K, const float* A1, int32_t lda1,
                       const float* B1, int32_t ldb1, const float* A2, int32_t lda2, const float* B2, int32_t ldb2,
                       float* C, int32_t ldc, int32_t accumulate);

/* The fused highway layer (north-star op):  h = act(S.Wh + bh), t = sigmoid(X.Wt + bt),
 * Y = t*h + (1-t)*X, with S = A_hat.X computed beforehand by gcnb_spmm_csr_f32.
 * highway_dense + MultiplicativeGatingLayer: gcnmodel.py:266,281-288.  Wh/Wt are (in, out)
 * row-major like Lasagne's DenseLayer.W.  H and T (saved for backward) may be NULL. */
int gcnb_highway_fwd_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t hd, const float* S, int32_t lds,
                         const float* X, int32_t ldx, const float* Wh, int32_t ldwh,
                         const float* bh, const float* Wt, int32_t ldwt, const float* bt,
                         int32_t act, float* Y, int32_t ldy, float* H, int32_t ldh, float* T,
                         int32_t ldt);
size_t gcnb_highway_workspace_bytes(int32_t n_rows, int32_t hd);

/* Y = T*H + (1-T)*X   (MultiplicativeGatingLayer.get_output_for, gcnmodel.py:266), element-wise */
int gcnb_highway_mix_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t hd, const float* H, int32_t ldh, const float* T,
                         int32_t ldt, const float* X, int32_t ldx, float* Y, int32_t ldy);

/* backward of the gate mix: dHpre = dY*T*act'(H), dTpre = dY*(H-X)*T*(1-T), dX = dY*(1-T) */
int gcnb_highway_bwd_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t hd, int32_t ld, const float* dY,
                         const float* X, const float* H, const float* T, int32_t act,
                         float* dHpre, float* dTpre, float* dX);

/* the same plus the two bias gradients dbh[k] = sum_rows dHpre[:, k], dbt[k] = sum_rows dTpre[:, k] (deterministic),
 * fused so that dHpre / dTpre are not read again for gcnb_colsum_f32.  Workspace: 2 x gcnb_colsum_workspace_bytes. */
int gcnb_highway_bwd_bias_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t hd, int32_t ld, const float* dY,
                              const float* X, const float* H, const float* T, int32_t act,
                              float* dHpre, float* dTpre, float* dX, float* dbh, float* dbt);

/* dZ = dY * keep*scale * act'(a), where Yact holds dropout(act(z)) (first layer,
 * gcnmodel.py:353-357) or act(z) when dropout_p == 0.  In-place (dZ == dY) allowed. */
int gcnb_act_bwd_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t k, int32_t ld, const float* dY,
                     const float* Yact, int32_t act, float dropout_p, uint64_t seed, int64_t row0,
                     float* dZ);

/* gcnb_act_bwd_f32 plus db[k] = sum_rows dZ[:, k] in the same pass (workspace: gcnb_colsum_workspace_bytes) */
int gcnb_act_bwd_bias_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t k, int32_t ld, const float* dY,
                          const float* Yact, int32_t act, float dropout_p, uint64_t seed, int64_t row0,
                          float* dZ, float* db);

/* out[k] (+)= sum over rows of A[:, k]  (bias gradients).  Deterministic. */
int gcnb_colsum_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t k, const float* A, int32_t lda,
                    float* out, int32_t accumulate);
size_t gcnb_colsum_workspace_bytes(int32_t n_rows, int32_t k);

/* dense[n_rows x n_cols] (leading dimension ld) = the CSR matrix, zeros elsewhere.  The dense hot-column block of X
 * travels host->device as CSR (a quarter of the bytes) and is expanded here.  Column ids must be < n_cols and
 * unique inside a row. */
int gcnb_csr_to_dense_f32(gcnb_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const float* val,
                          int32_t n_rows, int32_t n_cols, float* dense, int32_t ld);

/* dst[i, :k] = src[idx[i], :k] for i < n_idx (rows of a weight matrix selected by column id: the dense
 * hot-column block of X multiplies W0[hot, :]).  scatter: dst[idx[i], :k] = src[i, :k]. */
int gcnb_gather_rows_f32(gcnb_ctx* ctx, const float* src, int32_t ld_src, const int32_t* idx, int32_t n_idx,
                         int32_t k, float* dst, int32_t ld_dst);
int gcnb_scatter_rows_f32(gcnb_ctx* ctx, const float* src, int32_t ld_src, const int32_t* idx, int32_t n_idx,
                          int32_t k, float* dst, int32_t ld_dst);

/* ---------------------------------------------------------------- loss / outputs ------ */
/* metrics[0] += sum_i -log P[idx[i], labels[i]];  metrics[1] += #(argmax P[idx[i]] == labels[i])
 * categorical_crossentropy + argmax/eq: gcnmodel.py:376-389. */
int gcnb_xent_metrics_f32(gcnb_ctx* ctx, const float* P, int32_t ldp, int32_t n_classes,
                          const int32_t* idx, const int32_t* labels, int32_t n_idx,
                          float* metrics);
/* G = 0; G[idx[i], :] += (P[idx[i], :] - onehot(labels[i])) * inv_n   (d mean-CE / d logits) */
int gcnb_xent_grad_f32(gcnb_ctx* ctx, const float* P, int32_t ldp, int32_t n_classes,
                       int32_t n_rows, const int32_t* idx, const int32_t* labels, int32_t n_idx,
                       float inv_n, float* G, int32_t ldg);
/* the same gradient from a dense label map: row_label[r] (device int32, n_rows entries) = class of row r when r is a
 * training row, -1 otherwise (each training row counted once).  One pass writes every row of G; when a push is armed
 * (gcnb_push_arm) the rows also go to the owners of their columns for the graph convolution A^T.G that follows. */
int gcnb_xent_grad_dense_f32(gcnb_ctx* ctx, const float* P, int32_t ldp, int32_t n_classes, int32_t n_rows,
                             const int32_t* row_label, float inv_n, float* G, int32_t ldg);
/* preds[i] = argmax P[idx[i], :] (int64), probs[i, :] = P[idx[i], :] packed n_idx x n_classes
 * f_val outputs: gcnmodel.py:393-394,411. */
int gcnb_gather_argmax_f32(gcnb_ctx* ctx, const float* P, int32_t ldp, int32_t n_classes,
                           const int32_t* idx, int32_t n_idx, int64_t* preds, float* probs);

/* dist[i] = haversine km between (lat_true[i], lon_true[i]) and the median location of the predicted class
 * (class_lat[preds[i]], class_lon[preds[i]]): the loop of gcnmain.geo_eval (gcnmain.py:43-63) in float64.
 * `preds` is the int64 device array gcnb_gather_argmax_f32 fills, so predictions never leave the GPU.  *bad_flag
 * (device int32) is set when a prediction is outside [0, n_classes). */
int gcnb_geo_distance_f64(gcnb_ctx* ctx, const int64_t* preds, int32_t n, const double* class_lat,
                          const double* class_lon, int32_t n_classes, const double* lat_true,
                          const double* lon_true, double radius_km, double* dist, int32_t* bad_flag);

/* ---------------------------------------------------------------- optimiser ----------- */
/* G += coef*(sign(W) + 2W); reg_sum[0] += sum(|W| + W^2)   (gcnmodel.py:383-387) */
int gcnb_l1l2_f32(gcnb_ctx* ctx, const float* W, float* G, int64_t n, float coef, float* reg_sum);
/* One lasagne.updates.adam step (gcnmodel.py:407) over a flat parameter buffer.
 * state (device, 2 floats) = {t, a_t}; t is incremented on the device. */
int gcnb_adam_f32(gcnb_ctx* ctx, float* params, const float* grads, float* m, float* v, int64_t n,
                  float* state, float lr, float beta1, float beta2, float eps);

/* ---------------------------------------------------------------- dropout mask -------- */
/* materialise the keep mask the fused epilogue draws (tests feed it to the CPU oracle) */
int gcnb_dropout_mask_u8(gcnb_ctx* ctx, int32_t n_rows, int32_t k, float p, uint64_t seed,
                         int64_t row0, uint8_t* mask);

/* ---------------------------------------------------------------- A_hat construction -- */
/* The normalised adjacency the reference builds on the host before the hot path (gcnmain.py:115-128):
 *   adj = symmetric 0/1 adjacency of the undirected edge list (u[e], v[e]), duplicates merged;
 *   setdiag(0); setdiag(1); d = 1/sqrt(row sums); A_hat = (D * adj) * D in float64, cast to float32.
 * Output: CSR with ascending column indices inside each row (int32 / fp32, gcnmain.py:167-168).
 * Two calls because the caller owns all memory: build_rows fills rowptr[n_nodes + 1] (device) and returns
 * nnz to the host (one stream synchronisation); the caller allocates colidx / val of nnz entries and calls
 * fill with the same workspace.  `u`, `v` are device int32 arrays; self loops and repeated / reversed edges
 * are allowed (nx.Graph semantics); a node id outside [0, n_nodes) is GCNB_E_INVALID. */
size_t gcnb_adj_workspace_bytes(int64_t n_edges, int32_t n_nodes);
int gcnb_adj_build_rows(gcnb_ctx* ctx, const int32_t* u, const int32_t* v, int64_t n_edges, int32_t n_nodes,
                        void* work, size_t work_bytes, int32_t* rowptr, int64_t* nnz_host);
int gcnb_adj_fill_f32(gcnb_ctx* ctx, int64_t n_edges, int32_t n_nodes, const void* work, const int32_t* rowptr,
                      int32_t* colidx, float* val);

/* ---------------------------------------------------------------- multi-GPU exchange -- */
/* Row-partitioned runs on one NVSwitch box (the reference is single-process, gcnmodel.py:429-430).
 * Every rank (one process per GPU) owns an identically laid out ARENA of device memory that its peers map through
 * CUDA IPC; a graph convolution then runs feature-sliced: A_hat is replicated, rank q multiplies all rows of A_hat
 * by ITS column slice of the dense operand, and the two transposes around that product are stores into peer
 * memory (csrc/peer.cu).  Per rank and product N*K*4*(P-1)/P^2 bytes cross NVLink each way, against N*K*4*(P-1)/P
 * for an all-gather of the operand.
 *
 * The arena is the one allocation the library makes itself (allocator blocks cannot be exported through IPC):
 *   gcnb_peer_alloc  cudaMalloc + zero fill + IPC handle (GCNB_IPC_HANDLE_BYTES bytes, to be sent to the peers)
 *   gcnb_peer_open   map a peer's handle; gcnb_peer_close / gcnb_peer_free undo the two
 *   gcnb_peer_setup  register {base address of every rank's arena as mapped HERE, own rank included}, the arena
 *                    size and the offset of a GCNB_PEER_FLAG_BYTES flag block (zeroed) used by the barrier;
 *                    world == 0 detaches. */
#define GCNB_MAX_PEERS 16
#define GCNB_IPC_HANDLE_BYTES 64
#define GCNB_PEER_FLAG_BYTES 256
int gcnb_peer_alloc(gcnb_ctx* ctx, size_t bytes, void** dev_ptr, void* handle_out);
int gcnb_peer_free(gcnb_ctx* ctx, void* dev_ptr);
int gcnb_peer_open(gcnb_ctx* ctx, const void* handle, void** peer_ptr);
int gcnb_peer_close(gcnb_ctx* ctx, void* peer_ptr);
int gcnb_peer_setup(gcnb_ctx* ctx, int32_t rank, int32_t world, void* const* arena_base, size_t arena_bytes,
                    size_t flags_offset);
/* All ranks' streams meet: returns (on the stream) once every rank has launched the same barrier and everything
 * the ranks stored into peer memory before it is visible.  A rank that never arrives trips "peer_timeout_s"
 * (option, default 30) and the kernel traps. */
int gcnb_peer_barrier(gcnb_ctx* ctx);
/* rows -> column slices: copy x[0:n_loc, col0[q] : col0[q] + width[q]] (this rank's rows, global row0 + r) into rank
 * q's panel buffer at rows row0 + r, for every q (own rank included).  `xp_local` is THIS rank's address of the
 * panel buffer inside the arena (peers' addresses follow from the arena bases); ldp[q] is the leading dimension of
 * q's panel buffer.  col0 / width / ldp: host arrays of world entries, multiples of 4. */
int gcnb_slice_push_f32(gcnb_ctx* ctx, const float* x, int32_t ldx, int32_t n_loc, int64_t row0, float* xp_local,
                        const int32_t* col0, const int32_t* width, const int32_t* ldp);
/* Fused push: arm the context with a slice plan (same arguments as gcnb_slice_push_f32; K = operand width).  The NEXT
 * producer call whose output has K columns -- gcnb_spmm_csr_f32 (C, besides the local store), gcnb_highway_bwd_bias_f32
 * (dHpre, INSTEAD of the local store), gcnb_act_bwd_bias_f32 (dZ, besides), gcnb_xent_grad_dense_f32 (G, besides) --
 * stores that output straight into the owners' panel buffers from its own epilogue, so the transfer rides under the
 * producer's HBM traffic and gcnb_slice_push_f32 is not needed.  gcnb_push_consumed returns 1 if a producer did so
 * since arming (0: call gcnb_slice_push_f32) and disarms.  GCNB_E_UNSUPPORTED for K > 1024. */
int gcnb_push_arm(gcnb_ctx* ctx, float* xp_local, int32_t K, const int32_t* col0, const int32_t* width,
                  const int32_t* ldp, int64_t row0);
int gcnb_push_consumed(gcnb_ctx* ctx);

/* The sliced product: C[:, col0 : col0 + width] = epilogue(A . XP[:, 0 : width]) for ALL rows of A (the replicated
 * A_hat, n_rows = N); row i is stored into rank (i / n_pad)'s copy of C at local row i % n_pad.  `C` is this rank's
 * address of the output inside the arena (ldc floats per row, K = logical width of the whole operand for the
 * zero-padding rule), `bias` is indexed by the global column.  Epilogue: bias + activation only.  structured_dot(A, .)
 * of gcnmodel.py:130,153 and its gradient, each row summed in CSR order exactly as in gcnb_spmm_csr_f32. */
int gcnb_spmm_csr_sliced_f32(gcnb_ctx* ctx, const gcnb_csr* A, const float* XP, int32_t ldp, float* C, int32_t ldc,
                             int32_t K, int32_t col0, int32_t width, int32_t n_pad, const gcnb_epilogue* epi);
/* row softmax in place over C[n_rows x K] (after a sliced product + bias has landed); optional copy of the logits */
int gcnb_row_softmax_f32(gcnb_ctx* ctx, float* C, int32_t ldc, int32_t n_rows, int32_t K, float* logits);

/* Weighted graphs (nx.adjacency_matrix(..., weight='w'), gcnmain.py:115): `rowptr` / `colidx` / `weights` (device,
 * float64 like the SciPy matrix networkx returns) hold the symmetric weighted adjacency with its unit diagonal
 * already in place (gcnmain.py:117-120); val[k] = float32((d_i * w_k) * d_j), d = 1/sqrt(row sum) in float64, inf -> 0
 * (gcnmain.py:121-128).  `dinv_work`: n_nodes doubles of scratch. */
int gcnb_adj_normalize_weighted_f64(gcnb_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* weights,
                                    int32_t n_nodes, double* dinv_work, float* val);

#ifdef __cplusplus
}
#endif
#endif /* GCNB200_H */
