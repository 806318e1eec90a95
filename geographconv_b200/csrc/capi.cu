// capi.cu -- context management, transfers and the dense-op dispatchers of the C ABI
// (include/gcnb200.h).  The sparse and element-wise entry points live beside their kernels.
#include "common.cuh"

// implemented in gemm_simt.cu / elementwise.cu / gemm_tc.cu
int gcnb_gemm_simt(gcnb_ctx* ctx, int transA, int transB, int M, int N, int K, const float* A, int lda,
                   const float* B, int ldb, float* C, int ldc, int accumulate, const float* bias, int act);
int gcnb_highway_mix(gcnb_ctx* ctx, int n_rows, int hd, const float* H, int ldh, const float* T, int ldt,
                     const float* X, int ldx, float* Y, int ldy);
int gcnb_highway_tc(gcnb_ctx* ctx, int n_rows, int hd, const float* S, int lds, const float* X, int ldx,
                    const float* Wh, int ldwh, const float* bh, const float* Wt, int ldwt, const float* bt, int act,
                    float* Y, int ldy, float* H, int ldh, float* T, int ldt);
bool gcnb_highway_tc_supported(const gcnb_ctx* ctx, int n_rows, int hd, int lds, int ldx, int ldwh, int ldwt);
size_t gcnb_highway_tc_workspace_bytes(int hd);
int gcnb_gemm_tc(gcnb_ctx* ctx, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                 float* C, int ldc, const float* bias, int act, int accumulate);
bool gcnb_gemm_tc_supported(const gcnb_ctx* ctx, int transA, int transB, int M, int N, int K, int lda, int ldb,
                            int ldc, int accumulate);
size_t gcnb_gemm_tc_workspace_bytes(int N, int K);
int gcnb_gemm_pair_tc(gcnb_ctx* ctx, int transB, int M, int N, int K, const float* A1, int lda1, const float* B1, int ldb1,
                      const float* A2, int lda2, const float* B2, int ldb2, float* C, int ldc, int accumulate);
int gcnb_wgrad_tc(gcnb_ctx* ctx, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                  int ldc, int accumulate);
bool gcnb_wgrad_tc_supported(const gcnb_ctx* ctx, int M, int N, int K, int lda, int ldb);

extern "C" int gcnb_version(void) { return GCNB_VERSION; }

extern "C" int gcnb_create(int device, void* stream, gcnb_ctx** out) {
  if (!out) return GCNB_E_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return GCNB_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return GCNB_E_CUDA;
  gcnb_ctx* ctx = new gcnb_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return GCNB_E_CUDA;
  }
  if (prop.major != 10) {  // sm_100a SASS only: refuse to pretend on anything else
    delete ctx;
    return GCNB_E_UNSUPPORTED;
  }
  ctx->sm_count = prop.multiProcessorCount;
  if (stream) {
    ctx->stream = reinterpret_cast<cudaStream_t>(stream);
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete ctx;
      return GCNB_E_CUDA;
    }
    ctx->own_stream = true;
  }
  *out = ctx;
  return GCNB_OK;
}

extern "C" int gcnb_destroy(gcnb_ctx* ctx) {
  if (!ctx) return GCNB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& pp : ctx->pending) {
    cudaEventDestroy(pp.a);
    cudaEventDestroy(pp.b);
  }
  for (auto e : ctx->pool) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return GCNB_OK;
}

extern "C" const char* gcnb_last_error(const gcnb_ctx* ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int gcnb_set_stream(gcnb_ctx* ctx, void* stream) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, !ctx->prof || ctx->pending.empty(), "collect profiling events before switching streams");
  if (ctx->own_stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
    ctx->own_stream = false;
  }
  ctx->stream = reinterpret_cast<cudaStream_t>(stream);
  return GCNB_OK;
}
extern "C" void* gcnb_get_stream(const gcnb_ctx* ctx) { return ctx ? ctx->stream : nullptr; }

extern "C" int gcnb_set_workspace(gcnb_ctx* ctx, void* dev_ptr, size_t bytes) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, (dev_ptr != nullptr) == (bytes > 0), "workspace pointer and size disagree");
  GCNB_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(dev_ptr) & 255) == 0, "workspace must be 256-byte aligned");
  ctx->ws = dev_ptr;
  ctx->ws_bytes = bytes;
  return GCNB_OK;
}

static int* option_slot(gcnb_ctx* ctx, const char* name) {
  if (!name) return nullptr;
  if (!strcmp(name, "spmm_variant")) return &ctx->spmm_variant;
  if (!strcmp(name, "spmm_unroll")) return &ctx->spmm_unroll;
  if (!strcmp(name, "gemm_tc")) return &ctx->gemm_tc;
  if (!strcmp(name, "gemm_blo")) return &ctx->gemm_blo;
  if (!strcmp(name, "gemm_v")) return &ctx->gemm_v;
  if (!strcmp(name, "gemm_prefetch")) return &ctx->gemm_prefetch;
  if (!strcmp(name, "gemm_blo2")) return &ctx->gemm_blo2;
  if (!strcmp(name, "sm_margin")) return &ctx->sm_margin;
  if (!strcmp(name, "spmm_panel")) return &ctx->spmm_panel;
  if (!strcmp(name, "spmm_sliced_engine")) return &ctx->spmm_sliced_engine;
  if (!strcmp(name, "spmm_panel_policy")) return &ctx->spmm_panel_policy;
  if (!strcmp(name, "peer_timeout_s")) return &ctx->peer_timeout_s;
  if (!strcmp(name, "prof_mask")) return reinterpret_cast<int*>(&ctx->prof_mask);
  if (!strcmp(name, "tc_dbg_mode")) return &ctx->tc_dbg_mode;
  if (!strcmp(name, "tc_launches")) return &ctx->tc_launches;
  return nullptr;
}
extern "C" int gcnb_set_option(gcnb_ctx* ctx, const char* name, int value) {
  if (!ctx) return GCNB_E_INVALID;
  int* s = option_slot(ctx, name);
  if (!s) return gcnb_fail(ctx, GCNB_E_INVALID, "unknown option '%s'", name ? name : "(null)");
  *s = value;
  return GCNB_OK;
}
extern "C" int gcnb_get_option(const gcnb_ctx* ctx, const char* name, int* value) {
  if (!ctx || !value) return GCNB_E_INVALID;
  int* s = option_slot(const_cast<gcnb_ctx*>(ctx), name);
  if (!s) return GCNB_E_INVALID;
  *value = *s;
  return GCNB_OK;
}

// debug hook (not part of the public header): per-CTA clock64 phase stamps of the fused highway kernel
extern "C" int gcnb_debug_set_tc_buffer(gcnb_ctx* ctx, void* dev_ptr) {
  if (!ctx) return GCNB_E_INVALID;
  ctx->tc_dbg = dev_ptr;
  return GCNB_OK;
}

extern "C" int gcnb_sync(gcnb_ctx* ctx) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GCNB_OK;
}
extern "C" int gcnb_sm_count(const gcnb_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" long long gcnb_launch_count(const gcnb_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int gcnb_prof_enable(gcnb_ctx* ctx, int on) {
  if (!ctx) return GCNB_E_INVALID;
  ctx->prof = on != 0;
  return GCNB_OK;
}
static int prof_drain(gcnb_ctx* ctx) {
  GCNB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& pp : ctx->pending) {
    float ms = 0.f;
    GCNB_CUDA(ctx, cudaEventElapsedTime(&ms, pp.a, pp.b));
    ctx->prof_ms[pp.tag] += ms;
    ctx->prof_ops[pp.tag] += 1;
    ctx->pool.push_back(pp.a);
    ctx->pool.push_back(pp.b);
  }
  ctx->pending.clear();
  return GCNB_OK;
}
extern "C" int gcnb_prof_reset(gcnb_ctx* ctx) {
  if (!ctx) return GCNB_E_INVALID;
  int rc = prof_drain(ctx);
  for (int i = 0; i < GCNB_NTAGS; ++i) {
    ctx->prof_ms[i] = 0.f;
    ctx->prof_ops[i] = 0;
  }
  return rc;
}
extern "C" int gcnb_prof_collect(gcnb_ctx* ctx, float* ms, long long* ops) {
  if (!ctx) return GCNB_E_INVALID;
  int rc = prof_drain(ctx);
  if (rc != GCNB_OK) return rc;
  for (int i = 0; i < GCNB_NTAGS; ++i) {
    if (ms) ms[i] = ctx->prof_ms[i];
    if (ops) ops[i] = ctx->prof_ops[i];
  }
  return GCNB_OK;
}

extern "C" int gcnb_h2d(gcnb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  if (!ctx) return GCNB_E_INVALID;
  if (bytes == 0) return GCNB_OK;
  GCNB_REQUIRE(ctx, dst_dev && src_host, "null pointer");
  ProfScope scope(ctx, GCNB_TAG_COPY);
  GCNB_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return GCNB_OK;
}
extern "C" int gcnb_d2h(gcnb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  if (!ctx) return GCNB_E_INVALID;
  if (bytes == 0) return GCNB_OK;
  GCNB_REQUIRE(ctx, dst_host && src_dev, "null pointer");
  ProfScope scope(ctx, GCNB_TAG_COPY);
  GCNB_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return GCNB_OK;
}
extern "C" int gcnb_memset(gcnb_ctx* ctx, void* dst_dev, int byte, size_t bytes) {
  if (!ctx) return GCNB_E_INVALID;
  if (bytes == 0) return GCNB_OK;
  GCNB_REQUIRE(ctx, dst_dev, "null pointer");
  GCNB_CUDA(ctx, cudaMemsetAsync(dst_dev, byte, bytes, ctx->stream));
  return GCNB_OK;
}

extern "C" int gcnb_copy2d_f32(gcnb_ctx* ctx, const float* src, int32_t ld_src, float* dst, int32_t ld_dst,
                               int32_t rows, int32_t cols) {
  if (!ctx) return GCNB_E_INVALID;
  if (rows == 0 || cols == 0) return GCNB_OK;
  GCNB_REQUIRE(ctx, src && dst && ld_src >= cols && ld_dst >= cols, "bad 2-D copy");
  ProfScope scope(ctx, GCNB_TAG_COPY);
  GCNB_CUDA(ctx, cudaMemcpy2DAsync(dst, (size_t)ld_dst * sizeof(float), src, (size_t)ld_src * sizeof(float),
                                   (size_t)cols * sizeof(float), (size_t)rows, cudaMemcpyDeviceToDevice, ctx->stream));
  return GCNB_OK;
}

// ------------------------------------------------------------------------------- dense ops
extern "C" int gcnb_gemm_f32(gcnb_ctx* ctx, int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K,
                             const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
                             int32_t accumulate, const float* bias, int32_t act) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, A && B && C, "null matrix");
  GCNB_REQUIRE(ctx, M >= 0 && N > 0 && K > 0, "bad shape");
  GCNB_REQUIRE(ctx, lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "leading dimension too small");
  if (M == 0) return GCNB_OK;
  ProfScope scope(ctx, GCNB_TAG_GEMM);
  if (ctx->gemm_tc && transA && !transB && !bias && act == GCNB_ACT_LINEAR &&
      gcnb_wgrad_tc_supported(ctx, M, N, K, lda, ldb))
    return gcnb_wgrad_tc(ctx, M, N, K, A, lda, B, ldb, C, ldc, accumulate);
  if (ctx->gemm_tc && act <= GCNB_ACT_SIGMOID && gcnb_gemm_tc_supported(ctx, transA, transB, M, N, K, lda, ldb, ldc, accumulate))
    return gcnb_gemm_tc(ctx, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate);
  return gcnb_gemm_simt(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, accumulate, bias, act);
}

extern "C" int gcnb_gemm_pair_f32(gcnb_ctx* ctx, int32_t transB, int32_t M, int32_t N, int32_t K, const float* A1,
                                  int32_t lda1, const float* B1, int32_t ldb1, const float* A2, int32_t lda2,
                                  const float* B2, int32_t ldb2, float* C, int32_t ldc, int32_t accumulate) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, A1 && B1 && A2 && B2 && C, "null matrix");
  GCNB_REQUIRE(ctx, M >= 0 && N > 0 && K > 0, "bad shape");
  GCNB_REQUIRE(ctx, lda1 >= K && lda2 >= K && ldb1 >= (transB ? K : N) && ldb2 >= (transB ? K : N) && ldc >= N,
               "leading dimension too small");
  if (M == 0) return GCNB_OK;
  ProfScope scope(ctx, GCNB_TAG_GEMM);
  if (ctx->gemm_tc && gcnb_gemm_tc_supported(ctx, 0, transB, M, N, K, lda1, ldb1, ldc, accumulate) &&
      gcnb_gemm_tc_supported(ctx, 0, transB, M, N, K, lda2, ldb2, ldc, accumulate))
    return gcnb_gemm_pair_tc(ctx, transB, M, N, K, A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, C, ldc, accumulate);
  int rc = gcnb_gemm_simt(ctx, 0, transB, M, N, K, A1, lda1, B1, ldb1, C, ldc, accumulate, nullptr, GCNB_ACT_LINEAR);
  if (rc != GCNB_OK) return rc;
  return gcnb_gemm_simt(ctx, 0, transB, M, N, K, A2, lda2, B2, ldb2, C, ldc, 1, nullptr, GCNB_ACT_LINEAR);
}

extern "C" int gcnb_highway_mix_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t hd, const float* H, int32_t ldh, const float* T,
                                    int32_t ldt, const float* X, int32_t ldx, float* Y, int32_t ldy) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, H && T && X && Y, "null pointer");
  const int hd4 = ((hd + 3) / 4) * 4;
  GCNB_REQUIRE(ctx, ldh % 4 == 0 && ldt % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ldh >= hd4 && ldt >= hd4 &&
                        ldx >= hd4 && ldy >= hd4,
               "leading dimensions: multiple of 4, >= hd rounded to 4");
  if (n_rows == 0) return GCNB_OK;
  ProfScope scope(ctx, GCNB_TAG_ELEM);
  return gcnb_highway_mix(ctx, n_rows, hd, H, ldh, T, ldt, X, ldx, Y, ldy);
}

extern "C" size_t gcnb_highway_workspace_bytes(int32_t n_rows, int32_t hd) {
  // CUDA-core path: H and T scratch when the caller does not keep them; tcgen05 path: split weights
  const size_t ld = ((size_t)hd + 31) / 32 * 32;
  const size_t simt = 2 * (size_t)n_rows * ld * sizeof(float);
  const size_t tc = gcnb_highway_tc_workspace_bytes(hd);
  return simt > tc ? simt : tc;
}

extern "C" int gcnb_highway_fwd_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t hd, const float* S, int32_t lds,
                                    const float* X, int32_t ldx, const float* Wh, int32_t ldwh, const float* bh,
                                    const float* Wt, int32_t ldwt, const float* bt, int32_t act, float* Y,
                                    int32_t ldy, float* H, int32_t ldh, float* T, int32_t ldt) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, S && X && Wh && Wt && bh && bt && Y, "null pointer");
  const int hd4 = ((hd + 3) / 4) * 4;
  GCNB_REQUIRE(ctx, lds % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && lds >= hd4 && ldx >= hd4 && ldy >= hd4,
               "lds/ldx/ldy: multiple of 4, >= hd rounded to 4");
  GCNB_REQUIRE(ctx, (!H || (ldh % 4 == 0 && ldh >= hd4)) && (!T || (ldt % 4 == 0 && ldt >= hd4)), "ldh/ldt");
  if (n_rows == 0) return GCNB_OK;
  if (ctx->gemm_tc && act <= GCNB_ACT_SIGMOID && gcnb_highway_tc_supported(ctx, n_rows, hd, lds, ldx, ldwh, ldwt)) {
    ProfScope scope(ctx, GCNB_TAG_GEMM);
    return gcnb_highway_tc(ctx, n_rows, hd, S, lds, X, ldx, Wh, ldwh, bh, Wt, ldwt, bt, act, Y, ldy, H, ldh, T, ldt);
  }
  // CUDA-core composition: two GEMMs with fused bias+activation, then the gate mix
  float* Hb = H;
  float* Tb = T;
  int ldhb = ldh, ldtb = ldt;
  if (!Hb || !Tb) {
    const size_t ld = ((size_t)hd + 31) / 32 * 32;
    const size_t need = 2 * (size_t)n_rows * ld * sizeof(float);
    if (!ctx->ws || ctx->ws_bytes < need)
      return gcnb_fail(ctx, GCNB_E_WORKSPACE, "highway_fwd needs %s%lld workspace bytes, have %lld", "",
                       (long long)need, (long long)ctx->ws_bytes);
    if (!Hb) { Hb = reinterpret_cast<float*>(ctx->ws); ldhb = (int)ld; }
    if (!Tb) { Tb = reinterpret_cast<float*>(ctx->ws) + (size_t)n_rows * ld; ldtb = (int)ld; }
  }
  int rc;
  {
    ProfScope scope(ctx, GCNB_TAG_GEMM);
    rc = gcnb_gemm_simt(ctx, 0, 0, n_rows, hd, hd, S, lds, Wh, ldwh, Hb, ldhb, 0, bh, act);
    if (rc != GCNB_OK) return rc;
    rc = gcnb_gemm_simt(ctx, 0, 0, n_rows, hd, hd, X, ldx, Wt, ldwt, Tb, ldtb, 0, bt, GCNB_ACT_SIGMOID);
    if (rc != GCNB_OK) return rc;
  }
  ProfScope scope(ctx, GCNB_TAG_ELEM);
  return gcnb_highway_mix(ctx, n_rows, hd, Hb, ldhb, Tb, ldtb, X, ldx, Y, ldy);
}
