// common.cuh -- context, error plumbing, profiling scopes and small device helpers shared by
// every translation unit of libgcnb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/gcnb200.h"

struct ProfPair {
  cudaEvent_t a, b;
  int tag;
};

// Fused "rows -> column slices" transpose of the feature-sliced exchange (peer.cu): a producer kernel that is armed with
// a PushPlan stores its output row block straight into the owners' panel buffers over NVLink, instead of (or besides)
// writing it locally for gcnb_slice_push_f32 to copy.  Passed by value as a kernel parameter.
constexpr int kPushUnit = 16;       // slices are runs of 16-column units (partition.slice_columns)
constexpr int kPushMaxUnits = 64;   // operands up to 1024 columns
struct PushPlan {
  float* xp[GCNB_MAX_PEERS];   // rank q's panel buffer
  int col0[GCNB_MAX_PEERS];    // first column of q's slice
  int ldp[GCNB_MAX_PEERS];     // leading dimension of q's panel buffer
  unsigned char owner[kPushMaxUnits];  // owner of every 16-column unit
  long long row0;              // global index of local row 0
  int k4;                      // operand width rounded up to 4
  int on;
};

struct gcnb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  long long launches = 0;
  char err[512] = {0};
  // options
  int spmm_variant = 0;
  int spmm_unroll = 0;  // 0 = auto
  int sm_margin = 0;    // SMs the persistent SpMM kernel leaves free (for concurrently running NCCL kernels)
  int spmm_panel = 32;  // column-panel width (floats: 16, 32, 64) of the L2-resident panel engine (engine 2)
  int spmm_sliced_engine = -1;  // gather engine of the feature-sliced product: -1 by operand size, else 0 / 1 / 2
  int spmm_panel_policy = 1;  // gathers of the panel engine: 0 default policy, 1 L2 evict_last hint, 2 + L1 allocation
  int gemm_tc = 1;       // tcgen05 GEMMs where supported (0 = CUDA-core fp32 kernels only)
  int gemm_v = 1;        // tcgen05 GEMM kernel: 1 = one tile per SM (k-blocks of 32, default), 2 = two co-resident CTAs per SM
                         // (k-blocks of 16; 3-25% faster per product, parity-green, but its forward is bit-identical only among
                         // runs with the same GPU count class -- 1 GPU vs N GPUs differ in the last bit from the second highway
                         // layer on, cause not found: opt-in until it is)
  int gemm_prefetch = 0;   // gemm_v 2: L2 look-ahead of the activation rows in 128-row tiles (-1 = what the grid has in flight; measured
                           // slower than the plain demand loads: off)
  int gemm_blo2 = 0;       // gemm_v 2: weights' residual tile derived in shared memory (1) or loaded pre-split from L2 (0).  Measured
                           // (profiles/r2h_gemm_bench*.txt): 1 is 5% slower -- the kernel is bound by shared-memory bandwidth, and the
                           // extra converter traffic costs more than the TMA bytes it saves
  int gemm_blo = 0;      // 1: tcgen05 GEMMs derive the weights' lo tile in shared memory instead of loading it from L2
  int tc_launches = 0;
  int tc_dbg_mode = 0;
  void* tc_dbg = nullptr;  // device buffer for per-CTA phase timestamps of the highway kernel (debug tool)   // read-only: tcgen05 kernels launched so far
  // NVLink peer memory (peer.cu): identically laid out arenas of the ranks of one box
  char* peer_base[GCNB_MAX_PEERS] = {nullptr};
  int peer_rank = 0;
  int peer_world = 0;  // 0: no arena attached
  size_t peer_bytes = 0;
  size_t peer_flags_offset = 0;
  int peer_timeout_s = 30;
  PushPlan push_plan;          // armed by gcnb_push_arm, taken by the next producer that supports it
  bool push_armed = false;
  bool push_consumed = false;
  // profiling
  bool prof = false;
  unsigned prof_mask = 0xffffffffu;  // tags that are timed while profiling is on
  std::vector<ProfPair> pending;
  std::vector<cudaEvent_t> pool;
  float prof_ms[GCNB_NTAGS] = {0};
  long long prof_ops[GCNB_NTAGS] = {0};
};

static inline int gcnb_fail(gcnb_ctx* ctx, int code, const char* fmt, const char* a = "", long long b = 0,
                            long long c = 0) {
  if (ctx) snprintf(ctx->err, sizeof(ctx->err), fmt, a, b, c);
  return code;
}

#define GCNB_CUDA(ctx, call)                                                                   \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s -> %s", __FILE__, __LINE__,  \
                        #call, cudaGetErrorString(e__));                                       \
      return GCNB_E_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

#define GCNB_REQUIRE(ctx, cond, msg)                                                           \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d invalid: %s (%s)", __FILE__,    \
                        __LINE__, msg, #cond);                                                 \
      return GCNB_E_INVALID;                                                                   \
    }                                                                                          \
  } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define GCNB_LAUNCHED(ctx)                                                                     \
  do {                                                                                         \
    (ctx)->launches++;                                                                         \
    GCNB_CUDA(ctx, cudaGetLastError());                                                        \
  } while (0)

// RAII: CUDA-event pair around an op when profiling is on
struct ProfScope {
  gcnb_ctx* ctx;
  int tag;
  cudaEvent_t a = nullptr, b = nullptr;
  static cudaEvent_t get(gcnb_ctx* c) {
    if (!c->pool.empty()) {
      cudaEvent_t e = c->pool.back();
      c->pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  ProfScope(gcnb_ctx* c, int t) : ctx(c), tag(t) {
    if (ctx->prof && ((ctx->prof_mask >> t) & 1u)) {
      a = get(ctx);
      b = get(ctx);
      cudaEventRecord(a, ctx->stream);
    }
  }
  ~ProfScope() {
    if (a) {
      cudaEventRecord(b, ctx->stream);
      ctx->pending.push_back({a, b, tag});
    }
  }
};

// local arena pointer -> the same offset in rank q's arena; nullptr when [p, p + span) is not inside the local arena
void* gcnb_peer_translate(const gcnb_ctx* ctx, const void* p, int q, size_t span);

// the armed push plan if it matches an output of `k` columns (and disarm), else a plan with on == 0
static inline PushPlan gcnb_take_push(gcnb_ctx* ctx, int k) {
  PushPlan pp;
  memset(&pp, 0, sizeof(pp));
  if (ctx->push_armed && ctx->push_plan.k4 == ((k + 3) / 4) * 4) {
    pp = ctx->push_plan;
    pp.on = 1;
    ctx->push_armed = false;
    ctx->push_consumed = true;
  }
  return pp;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------- device helpers
__device__ __forceinline__ float act_apply(int act, float z) {
  switch (act) {
    case GCNB_ACT_TANH: return tanhf(z);
    case GCNB_ACT_RELU: return fmaxf(z, 0.f);
    case GCNB_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    case GCNB_ACT_SELU: return 1.0507009873554805f * (z > 0.f ? z : 1.6732632423543772f * expm1f(z));
    default: return z;
  }
}
// derivative wrt the pre-activation, expressed through the activation output
__device__ __forceinline__ float act_grad_from_out(int act, float y) {
  switch (act) {
    case GCNB_ACT_TANH: return 1.f - y * y;
    case GCNB_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case GCNB_ACT_SIGMOID: return y * (1.f - y);
    case GCNB_ACT_SELU: return y > 0.f ? 1.0507009873554805f : y + 1.0507009873554805f * 1.6732632423543772f;
    default: return 1.f;
  }
}

// Philox4x32-10 (Salmon et al. 2011): counter (c0..c3), key (k0,k1) -> 4 x uint32.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  double t = (1.0 - (double)p) * 4294967296.0;
  if (t > 4294967295.0) t = 4294967295.0;
  if (t < 0.0) t = 0.0;
  return (uint32_t)t;
}
// keep bits of the four columns 4*g .. 4*g+3 of global row `grow`
__device__ __forceinline__ uint4 dropout_draw(uint64_t seed, int64_t grow, uint32_t g) {
  return philox4x32_10(make_uint4((uint32_t)grow, g, 0u, 0u),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// streaming (read-once) loads: do not allocate in L1, evict-first in L2 (cache-policy operand form;
// ptxas only accepts the plain .L2::evict_first qualifier on 256-bit loads)
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ int ld_stream_s32(const int* p, uint64_t pol) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float ld_stream_f32(const float* p, uint64_t pol) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
  return r;
}
// gathered dense rows: no L1 allocation (no reuse inside an SM), normal L2 policy (rows are re-used
// across SMs, the 126 MB L2 serves a share of the gathers)
__device__ __forceinline__ float4 ld_gather_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// store one float4 (columns c .. c+3 of local row r) into the panel buffer of the rank that owns those columns
__device__ __forceinline__ void push_store_f4(const PushPlan& pp, long long r, int c, float4 v) {
  if (c >= pp.k4) return;
  const int q = pp.owner[c >> 4];
  float4* dst = reinterpret_cast<float4*>(pp.xp[q] + (size_t)(pp.row0 + r) * pp.ldp[q] + (c - pp.col0[q]));
  *dst = v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
