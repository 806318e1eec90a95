// peer.cu -- NVLink peer memory for row-partitioned runs on one NVSwitch box.
//
// The reference has no distributed code; the build adds ONE exchange per graph convolution.  The first design
// all-gathered the whole N x K dense operand into every rank (NCCL), which moves N*K*4 bytes into every GPU whatever
// the number of ranks.  The exchange here is the feature-sliced one: A_hat is replicated (it is small: nnz*8 bytes),
// every rank multiplies ALL rows of A_hat by ITS slice of the operand's columns, and the two transposes around that
// product (rows -> column slices, column slices -> rows) are plain stores into the peers' memory:
//
//   gcnb_slice_push_f32      x[my rows, cols of q]  ->  q's panel buffer XP_q[my rows, :]        (NVLink stores)
//   gcnb_peer_barrier        every rank's pushes have landed
//   gcnb_spmm_csr_sliced_f32 out[all rows, my cols] = A_hat . XP_me ; row i is stored into its owner's buffer (spmm.cu)
//   gcnb_peer_barrier        every rank's result columns have landed
//
// Per rank and product that is N*K*4*(P-1)/P^2 bytes each way instead of N*K*4*(P-1)/P, and each row is still summed
// in CSR order by one lane group, so the result is bit-identical to the single-GPU product.
//
// Peer memory is one cudaMalloc'ed arena per rank exported with CUDA IPC (the one place the library allocates device
// memory: the torch caching allocator's blocks cannot be exported); the ranks lay their arenas out identically, so a
// local pointer translates to peer q's copy by base-address arithmetic.
#include "common.cuh"

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct PeerTab {
  uint32_t* flags[GCNB_MAX_PEERS];  // flags[q]: rank q's flag block (GCNB_MAX_PEERS arrival words + the epoch word)
};

// One CTA, one thread per peer.  Epoch e of the barrier: thread q stores e into peer q's arrival word [rank], then
// spins until its own arrival word [q] reaches e.  Kernels earlier on the stream have completed, so their stores into
// peer memory are performed; the system-scope fence + release / acquire pair orders them before the peers' next reads.
// A peer that never arrives (its process died) trips the timeout and traps: the run fails loudly instead of hanging.
__global__ void peer_barrier_kernel(PeerTab tab, int rank, int world, unsigned long long timeout_ns) {
  __shared__ uint32_t epoch_s;
  uint32_t* mine = tab.flags[rank];
  if (threadIdx.x == 0) {
    epoch_s = mine[GCNB_MAX_PEERS] + 1u;
    mine[GCNB_MAX_PEERS] = epoch_s;
  }
  __syncthreads();
  const uint32_t epoch = epoch_s;
  const int q = threadIdx.x;
  if (q < world) {
    __threadfence_system();
    st_release_sys(tab.flags[q] + rank, epoch);
    const unsigned long long t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(mine + q) - epoch) < 0) {
      if (globaltimer_ns() - t0 > timeout_ns) {
        printf("gcnb peer barrier: rank %d waited too long for rank %d (epoch %u)\n", rank, q, epoch);
        __trap();
      }
    }
    __threadfence_system();
  }
}

struct PushParams {
  const float* x;
  int ldx;
  int n_loc;
  long long row0;                 // global index of local row 0
  float* xp[GCNB_MAX_PEERS];      // peer q's panel buffer (its own address of the same arena offset)
  int col0[GCNB_MAX_PEERS];       // first column of q's slice
  int nf4[GCNB_MAX_PEERS];        // float4s per row in q's slice
  int ldp[GCNB_MAX_PEERS];        // leading dimension of q's panel buffer
  int world;
  int rank;
};

// blockIdx.y = destination rank (rotated so that at any moment the ranks write to different peers); a thread moves one
// float4, consecutive threads consecutive float4s of a row slice: 128-byte (or longer) contiguous NVLink stores
__global__ void __launch_bounds__(256) slice_push_kernel(const PushParams p) {
  const int q = (p.rank + 1 + (int)blockIdx.y) % p.world;
  const int nf4 = p.nf4[q];
  if (nf4 == 0) return;
  const long long total = (long long)p.n_loc * nf4;
  float* dst = p.xp[q];
  const int ldp = p.ldp[q];
  const int c0 = p.col0[q];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nf4;
    const int c = (int)(i - r * nf4);
    const float4 v = *reinterpret_cast<const float4*>(p.x + (size_t)r * p.ldx + c0 + 4 * c);
    *reinterpret_cast<float4*>(dst + (size_t)(p.row0 + r) * ldp + 4 * c) = v;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ arena
extern "C" int gcnb_peer_alloc(gcnb_ctx* ctx, size_t bytes, void** dev_ptr, void* handle_out) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, dev_ptr && handle_out && bytes > 0, "null pointer / empty arena");
  GCNB_CUDA(ctx, cudaSetDevice(ctx->device));
  void* p = nullptr;
  GCNB_CUDA(ctx, cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemsetAsync(p, 0, bytes, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    snprintf(ctx->err, sizeof(ctx->err), "peer arena of %lld bytes: %s", (long long)bytes, cudaGetErrorString(e));
    return GCNB_E_CUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == GCNB_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return GCNB_OK;
}

extern "C" int gcnb_peer_free(gcnb_ctx* ctx, void* dev_ptr) {
  if (!ctx) return GCNB_E_INVALID;
  if (!dev_ptr) return GCNB_OK;
  GCNB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  GCNB_CUDA(ctx, cudaFree(dev_ptr));
  return GCNB_OK;
}

extern "C" int gcnb_peer_open(gcnb_ctx* ctx, const void* handle, void** peer_ptr) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, handle && peer_ptr, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  GCNB_CUDA(ctx, cudaSetDevice(ctx->device));
  GCNB_CUDA(ctx, cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return GCNB_OK;
}

extern "C" int gcnb_peer_close(gcnb_ctx* ctx, void* peer_ptr) {
  if (!ctx) return GCNB_E_INVALID;
  if (!peer_ptr) return GCNB_OK;
  GCNB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  GCNB_CUDA(ctx, cudaIpcCloseMemHandle(peer_ptr));
  return GCNB_OK;
}

extern "C" int gcnb_peer_setup(gcnb_ctx* ctx, int32_t rank, int32_t world, void* const* arena_base, size_t arena_bytes,
                               size_t flags_offset) {
  if (!ctx) return GCNB_E_INVALID;
  if (world == 0) {  // detach
    ctx->peer_world = 0;
    return GCNB_OK;
  }
  GCNB_REQUIRE(ctx, world >= 2 && world <= GCNB_MAX_PEERS && rank >= 0 && rank < world, "bad rank / world");
  GCNB_REQUIRE(ctx, arena_base && arena_bytes > 0, "null arena table");
  GCNB_REQUIRE(ctx, flags_offset % 256 == 0 && flags_offset + GCNB_PEER_FLAG_BYTES <= arena_bytes, "flag block outside the arena");
  for (int q = 0; q < world; ++q) {
    GCNB_REQUIRE(ctx, arena_base[q] != nullptr, "null peer arena");
    ctx->peer_base[q] = reinterpret_cast<char*>(arena_base[q]);
  }
  ctx->peer_rank = rank;
  ctx->peer_world = world;
  ctx->peer_bytes = arena_bytes;
  ctx->peer_flags_offset = flags_offset;
  return GCNB_OK;
}

// local arena pointer -> the same offset inside rank q's arena (nullptr when `p` is not inside the local arena)
void* gcnb_peer_translate(const gcnb_ctx* ctx, const void* p, int q, size_t span) {
  if (ctx->peer_world < 2) return nullptr;
  const char* base = ctx->peer_base[ctx->peer_rank];
  const char* c = reinterpret_cast<const char*>(p);
  if (c < base || c + span > base + ctx->peer_bytes) return nullptr;
  return ctx->peer_base[q] + (c - base);
}

extern "C" int gcnb_peer_barrier(gcnb_ctx* ctx) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, ctx->peer_world >= 2, "no peer arena attached (gcnb_peer_setup)");
  ProfScope scope(ctx, GCNB_TAG_SYNC);
  PeerTab tab;
  memset(&tab, 0, sizeof(tab));
  for (int q = 0; q < ctx->peer_world; ++q)
    tab.flags[q] = reinterpret_cast<uint32_t*>(ctx->peer_base[q] + ctx->peer_flags_offset);
  const unsigned long long timeout_ns = (unsigned long long)(ctx->peer_timeout_s > 0 ? ctx->peer_timeout_s : 30) * 1000000000ull;
  peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(tab, ctx->peer_rank, ctx->peer_world, timeout_ns);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" int gcnb_slice_push_f32(gcnb_ctx* ctx, const float* x, int32_t ldx, int32_t n_loc, int64_t row0,
                                   float* xp_local, const int32_t* col0, const int32_t* width, const int32_t* ldp) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, ctx->peer_world >= 2, "no peer arena attached (gcnb_peer_setup)");
  GCNB_REQUIRE(ctx, x && xp_local && col0 && width && ldp, "null pointer");
  GCNB_REQUIRE(ctx, aligned16(x) && aligned16(xp_local) && ldx % 4 == 0, "16-byte alignment");
  if (n_loc == 0) return GCNB_OK;
  PushParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.ldx = ldx; p.n_loc = n_loc; p.row0 = row0; p.world = ctx->peer_world; p.rank = ctx->peer_rank;
  int max_nf4 = 0;
  for (int q = 0; q < ctx->peer_world; ++q) {
    GCNB_REQUIRE(ctx, col0[q] % 4 == 0 && width[q] % 4 == 0 && ldp[q] % 4 == 0 && width[q] >= 0 && ldp[q] >= width[q] &&
                          col0[q] + width[q] <= ldx,
                 "slice: columns in multiples of 4 inside the operand");
    p.xp[q] = reinterpret_cast<float*>(gcnb_peer_translate(ctx, xp_local, q, 16));
    GCNB_REQUIRE(ctx, p.xp[q] != nullptr, "panel buffer is not inside the peer arena");
    p.col0[q] = col0[q]; p.nf4[q] = width[q] / 4; p.ldp[q] = ldp[q];
    if (p.nf4[q] > max_nf4) max_nf4 = p.nf4[q];
  }
  if (max_nf4 == 0) return GCNB_OK;
  ProfScope scope(ctx, GCNB_TAG_COMM);
  long long blocks = ((long long)n_loc * max_nf4 + 255) / 256;
  const long long cap = (long long)ctx->sm_count * 8 / ctx->peer_world + 1;
  if (blocks > cap) blocks = cap;
  const dim3 grid((unsigned)blocks, (unsigned)ctx->peer_world);
  slice_push_kernel<<<grid, 256, 0, ctx->stream>>>(p);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

// ------------------------------------------------------------------------------------------------ fused push
extern "C" int gcnb_push_arm(gcnb_ctx* ctx, float* xp_local, int32_t K, const int32_t* col0, const int32_t* width,
                             const int32_t* ldp, int64_t row0) {
  if (!ctx) return GCNB_E_INVALID;
  ctx->push_armed = false;
  ctx->push_consumed = false;
  GCNB_REQUIRE(ctx, ctx->peer_world >= 2, "no peer arena attached (gcnb_peer_setup)");
  GCNB_REQUIRE(ctx, xp_local && col0 && width && ldp && K > 0, "null pointer");
  const int k4 = ((K + 3) / 4) * 4;
  if ((k4 + kPushUnit - 1) / kPushUnit > kPushMaxUnits) return GCNB_E_UNSUPPORTED;
  PushPlan pp;
  memset(&pp, 0, sizeof(pp));
  pp.row0 = row0;
  pp.k4 = k4;
  int next = 0;
  for (int q = 0; q < ctx->peer_world; ++q) {
    GCNB_REQUIRE(ctx, col0[q] % 4 == 0 && width[q] % 4 == 0 && ldp[q] % 4 == 0 && width[q] >= 0 && ldp[q] >= width[q],
                 "slice: columns in multiples of 4");
    pp.xp[q] = reinterpret_cast<float*>(gcnb_peer_translate(ctx, xp_local, q, 16));
    GCNB_REQUIRE(ctx, pp.xp[q] != nullptr, "panel buffer is not inside the peer arena");
    pp.col0[q] = col0[q];
    pp.ldp[q] = ldp[q];
    if (width[q] == 0) continue;
    // slices are contiguous runs of whole 16-column units in rank order (the last one may be cut at k4)
    GCNB_REQUIRE(ctx, col0[q] == next && col0[q] % kPushUnit == 0, "slices must tile the columns in rank order, unit-aligned");
    for (int c = col0[q]; c < col0[q] + width[q]; c += kPushUnit) pp.owner[c / kPushUnit] = (unsigned char)q;
    next = col0[q] + ((width[q] + kPushUnit - 1) / kPushUnit) * kPushUnit;
  }
  GCNB_REQUIRE(ctx, next >= k4, "slices do not cover the operand");
  ctx->push_plan = pp;
  ctx->push_armed = true;
  return GCNB_OK;
}

extern "C" int gcnb_push_consumed(gcnb_ctx* ctx) {
  if (!ctx) return 0;
  const int done = ctx->push_consumed ? 1 : 0;
  ctx->push_armed = false;
  ctx->push_consumed = false;
  return done;
}
