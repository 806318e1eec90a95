// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMMs for sm_100a with fp32-grade accuracy.
//
// Replaces T.dot (reference gcnmodel.py:126,149,285) and its dgrad products on the hot path, and
// fuses the whole highway layer (gcnmodel.py:266,281-288) into one kernel:
//     C = epilogue(A . Bt^T)          A: M x K row-major, Bt: N x K row-major (both "K-major")
//     highway:  h = act(S.Wh + bh), t = sigmoid(X.Wt + bt), Y = t*h + (1-t)*X   (two accumulators)
//
// Precision: the reference computes in fp32 (BLAS sgemm).  kind::tf32 alone (10-bit mantissa) misses
// the 1e-3 parity budget after a few layers, so every product is done as an error-compensated
// 3xTF32 split: a = a_hi + a_lo with a_hi = rn_tf32(a), a_lo = a - a_hi (exact in fp32);
// A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, all accumulated in fp32 in TMEM (error ~2^-21).
//
// Structure (one 128 x BN output tile per CTA, 320 threads, 1 CTA / SM):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes (128B-swizzled) of raw fp32 A and Bt
//               tiles into a 3-stage shared-memory ring, completion on `full` mbarriers;
//   warps 2-9   converters: split every landed tile in place into hi (tf32-rounded) and a second
//               `lo` tile with the same swizzled layout, fence.proxy.async, arrive on `conv`;
//   warp 1      allocates TMEM, one elected lane issues 12 tcgen05.mma.kind::tf32 (M128 x BN x K8)
//               per 32-wide k-block, tcgen05.commit releases the stage (`empty`) and finally
//               signals `acc_full`;
//   warps 2-9   epilogue: tcgen05.ld the fp32 accumulators (32 lanes x 32 columns per warp),
//               bias / activation / gate mix / accumulate, 128-byte row segments to global.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;  // floats: 128 bytes = one SWIZZLE_128B atom row; 4 tf32 MMAs (K = 8) per k-block
constexpr int kStages = 3;
constexpr int kConvWarps = 8;
constexpr int kThreads = 64 + kConvWarps * 32;
constexpr int kTmemCols = 512;
constexpr int kMaxBN = 160;

struct alignas(64) TcParams {
  CUtensorMap mapA[2];
  CUtensorMap mapB[2];
  int nphase;      // 1: plain GEMM, 2: fused highway (phase 0 = S.Wh, phase 1 = X.Wt)
  int kblocks[2];  // 32-wide k-blocks per phase
  int M, N, BN, n_tiles;
  float* C;        // plain: output; highway: Y
  int ldc;
  const float* bias;  // plain: bias or null; highway: bh
  int act;
  int accumulate;
  // highway only
  const float* bias_t;
  const float* X;
  int ldx;
  float* H;
  int ldh;
  float* T;
  int ldt;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8-row groups
// 1024 B apart (SBO), LBO = 1 (unused for swizzled K-major), version 1, layout type 2.
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = bn
__device__ __forceinline__ uint32_t umma_idesc_tf32(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte alignment of every tile (SWIZZLE_128B atoms are 8 rows x 128 B)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  const uint32_t a_bytes = BM * 128, b_bytes = (uint32_t)BN * 128;
  const uint32_t half_bytes = a_bytes + b_bytes;   // [A_hi | B_hi] then [A_lo | B_lo]
  const uint32_t stage_bytes = 2 * half_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * stage_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_conv = bar_full + 8 * kStages, bar_empty = bar_conv + 8 * kStages;
  const uint32_t bar_acc = bar_empty + 8 * kStages;

  const int n_tile = blockIdx.x % p.n_tiles, m_tile = blockIdx.x / p.n_tiles;
  const int m0 = m_tile * BM, n0 = n_tile * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, kConvWarps);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int total_kb = p.kblocks[0] + (p.nphase > 1 ? p.kblocks[1] : 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int ph = 0; ph < p.nphase; ++ph) {
        for (int kb = 0; kb < p.kblocks[ph]; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t par = (it / kStages) & 1;
          mbar_wait(bar_empty + 8 * s, par ^ 1);
          mbar_expect_tx(bar_full + 8 * s, half_bytes);
          const uint32_t dst = smem_base + s * stage_bytes;
          tma_load_2d(dst, &p.mapA[ph], kb * BK, m0, bar_full + 8 * s);
          tma_load_2d(dst + a_bytes, &p.mapB[ph], kb * BK, n0, bar_full + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc_tf32(BN);
    int it = 0;
    for (int ph = 0; ph < p.nphase; ++ph) {
      const uint32_t tacc = tmem_base + (uint32_t)(ph * BN);
      for (int kb = 0; kb < p.kblocks[ph]; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t par = (it / kStages) & 1;
        mbar_wait(bar_conv + 8 * s, par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t a_hi = smem_base + s * stage_bytes, b_hi = a_hi + a_bytes;
          const uint32_t a_lo = a_hi + half_bytes, b_lo = b_hi + half_bytes;
          // small terms first, then hi.hi
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)
            umma_tf32(tacc, umma_desc_k128(a_lo + 32 * k), umma_desc_k128(b_hi + 32 * k), idesc, (kb | k) != 0);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)
            umma_tf32(tacc, umma_desc_k128(a_hi + 32 * k), umma_desc_k128(b_lo + 32 * k), idesc, 1u);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)
            umma_tf32(tacc, umma_desc_k128(a_hi + 32 * k), umma_desc_k128(b_hi + 32 * k), idesc, 1u);
          umma_commit(bar_empty + 8 * s);  // implies tcgen05.fence::before_thread_sync
          if (it == total_kb - 1) umma_commit(bar_acc);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ converters, then epilogue
    const int ct = threadIdx.x - 64;  // 0 .. 255
    const int n_chunks = (int)(half_bytes >> 4);
    for (int it = 0; it < total_kb; ++it) {
      const int s = it % kStages;
      const uint32_t par = (it / kStages) & 1;
      mbar_wait(bar_full + 8 * s, par);
      unsigned char* hi = smem + (size_t)s * stage_bytes;
      unsigned char* lo = hi + half_bytes;
      for (int c = ct; c < n_chunks; c += kConvWarps * 32) {
        float4 v = *reinterpret_cast<const float4*>(hi + 16 * c);
        float4 h, l;
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t); l.x = v.x - h.x;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t); l.y = v.y - h.y;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t); l.z = v.z - h.z;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t); l.w = v.w - h.w;
        *reinterpret_cast<float4*>(hi + 16 * c) = h;
        *reinterpret_cast<float4*>(lo + 16 * c) = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor core reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_conv + 8 * s);
    }

    mbar_wait(bar_acc, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // two warps share a quarter: interleave the 32-column chunks
    const int row = m0 + q * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool row_ok = row < p.M;
    for (int ch = half; ch < BN / 32; ch += 2) {
      const int col0 = n0 + ch * 32;
      if (col0 >= p.N) break;  // warp-uniform
      float acc[32];
      tmem_ld32(tlane + (uint32_t)(ch * 32), acc);
      if (p.nphase == 1) {
        float* crow = p.C + (size_t)row * p.ldc + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = col0 + j + e;
            float v = acc[j + e];
            if (!p.accumulate) {
              if (p.bias != nullptr && col < p.N) v += __ldg(p.bias + col);
              v = act_apply(p.act, v);
            }
            o[e] = v;
          }
          if (row_ok) {
            if (col0 + j + 3 < p.N) {
              float4 w = make_float4(o[0], o[1], o[2], o[3]);
              if (p.accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(crow + j);
                w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
              }
              *reinterpret_cast<float4*>(crow + j) = w;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (col0 + j + e < p.N) crow[j + e] = p.accumulate ? crow[j + e] + o[e] : o[e];
            }
          }
        }
      } else {
        float acc_t[32];
        tmem_ld32(tlane + (uint32_t)(BN + ch * 32), acc_t);
        const float* xrow = p.X + (size_t)row * p.ldx + col0;
        float* yrow = p.C + (size_t)row * p.ldc + col0;
        float* hrow = p.H ? p.H + (size_t)row * p.ldh + col0 : nullptr;
        float* trow = p.T ? p.T + (size_t)row * p.ldt + col0 : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float h[4], t[4], y[4], x[4] = {0.f, 0.f, 0.f, 0.f};
          const bool full4 = col0 + j + 3 < p.N;
          if (row_ok) {
            if (full4) {
              const float4 xv = *reinterpret_cast<const float4*>(xrow + j);
              x[0] = xv.x; x[1] = xv.y; x[2] = xv.z; x[3] = xv.w;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (col0 + j + e < p.N) x[e] = xrow[j + e];
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = col0 + j + e;
            const bool ok = col < p.N;
            h[e] = act_apply(p.act, acc[j + e] + (ok ? __ldg(p.bias + col) : 0.f));
            t[e] = sigmoidf_(acc_t[j + e] + (ok ? __ldg(p.bias_t + col) : 0.f));
            y[e] = t[e] * h[e] + (1.0f - t[e]) * x[e];
          }
          if (row_ok) {
            if (full4) {
              *reinterpret_cast<float4*>(yrow + j) = make_float4(y[0], y[1], y[2], y[3]);
              if (hrow) *reinterpret_cast<float4*>(hrow + j) = make_float4(h[0], h[1], h[2], h[3]);
              if (trow) *reinterpret_cast<float4*>(trow + j) = make_float4(t[0], t[1], t[2], t[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (col0 + j + e < p.N) {
                  yrow[j + e] = y[e];
                  if (hrow) hrow[j + e] = h[e];
                  if (trow) trow[j + e] = t[e];
                }
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad:  C[M x N] (+)= A^T . B with A: K x M and B: K x N row-major, K = number of graph nodes
// (dW = x^T . V, gcnmodel.py:407's gradient of T.dot wrt W).  Both operands are "MN-major" for the
// tensor core: a shared-memory tile holds 32-float column chunks, each [32 k-rows][128 B] with the
// 128-byte swizzle TMA writes; the descriptor's LBO is the distance between chunks and one
// tcgen05.mma consumes one 8-row k-group (1024 B).  The K range is split over CTAs (split-K); each
// CTA stores its raw 128 x BN tile to a partial buffer that splitk_reduce adds in split order, so
// the result does not depend on scheduling.
// ---------------------------------------------------------------------------------------------
struct alignas(64) WgParams {
  CUtensorMap mapA;  // [K rows][M cols], box 32 x 32
  CUtensorMap mapB;  // [K rows][N cols], box 32 x 32
  int M, N, K, BN, m_tiles, n_tiles, splits, kb_per_split;
  float* partial;    // [splits][M][N]
};

// MN-major descriptor for 32-bit operands: tf32 only accepts SWIZZLE_128B_BASE32B (layout type 1; 32-byte
// units XOR-ed over a 4-row period, what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
// LBO = distance between 32-column chunks, SBO = distance between 4-row k-groups (512 B).
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t saddr, uint32_t chunk_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(chunk_bytes >> 4) << 16) | (32ull << 32) | (1ull << 46) |
         (1ull << 61);
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  constexpr uint32_t chunk_bytes = BK * 128;  // one 32-column chunk: 32 k-rows x 128 B
  const uint32_t a_bytes = (BM / 32) * chunk_bytes, b_bytes = (uint32_t)(BN / 32) * chunk_bytes;
  const uint32_t half_bytes = a_bytes + b_bytes;
  const uint32_t stage_bytes = 2 * half_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * stage_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_conv = bar_full + 8 * kStages, bar_empty = bar_conv + 8 * kStages;
  const uint32_t bar_acc = bar_empty + 8 * kStages;

  const int tiles = p.m_tiles * p.n_tiles;
  const int tile = blockIdx.x % tiles, split = blockIdx.x / tiles;
  const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
  const int m0 = m_tile * BM, n0 = n_tile * BN;
  const int kb0 = split * p.kb_per_split;
  const int kb_total = (p.K + BK - 1) / BK;
  int nkb = kb_total - kb0;
  if (nkb > p.kb_per_split) nkb = p.kb_per_split;
  if (nkb < 0) nkb = 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, kConvWarps);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < nkb; ++it) {
        const int s = it % kStages;
        const uint32_t par = (it / kStages) & 1;
        mbar_wait(bar_empty + 8 * s, par ^ 1);
        mbar_expect_tx(bar_full + 8 * s, half_bytes);
        const uint32_t dst = smem_base + s * stage_bytes;
        const int k0 = (kb0 + it) * BK;
        for (int c = 0; c < BM / 32; ++c) tma_load_2d(dst + c * chunk_bytes, &p.mapA, m0 + 32 * c, k0, bar_full + 8 * s);
        for (int c = 0; c < BN / 32; ++c)
          tma_load_2d(dst + a_bytes + c * chunk_bytes, &p.mapB, n0 + 32 * c, k0, bar_full + 8 * s);
      }
    }
  } else if (warp == 1) {
    // MN-major A and B: bits 15 and 16 of the instruction descriptor
    const uint32_t idesc = umma_idesc_tf32(BN) | (1u << 15) | (1u << 16);
    for (int it = 0; it < nkb; ++it) {
      const int s = it % kStages;
      const uint32_t par = (it / kStages) & 1;
      mbar_wait(bar_conv + 8 * s, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_hi = smem_base + s * stage_bytes, b_hi = a_hi + a_bytes;
        const uint32_t a_lo = a_hi + half_bytes, b_lo = b_hi + half_bytes;
#pragma unroll
        for (int k = 0; k < BK / 8; ++k)
          umma_tf32(tmem_base, umma_desc_mn128(a_lo + 1024 * k, chunk_bytes), umma_desc_mn128(b_hi + 1024 * k, chunk_bytes), idesc, (it | k) != 0);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k)
          umma_tf32(tmem_base, umma_desc_mn128(a_hi + 1024 * k, chunk_bytes), umma_desc_mn128(b_lo + 1024 * k, chunk_bytes), idesc, 1u);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k)
          umma_tf32(tmem_base, umma_desc_mn128(a_hi + 1024 * k, chunk_bytes), umma_desc_mn128(b_hi + 1024 * k, chunk_bytes), idesc, 1u);
        umma_commit(bar_empty + 8 * s);
        if (it == nkb - 1) umma_commit(bar_acc);
      }
      __syncwarp();
    }
  } else {
    const int ct = threadIdx.x - 64;
    const int n_chunks = (int)(half_bytes >> 4);
    for (int it = 0; it < nkb; ++it) {
      const int s = it % kStages;
      const uint32_t par = (it / kStages) & 1;
      mbar_wait(bar_full + 8 * s, par);
      unsigned char* hi = smem + (size_t)s * stage_bytes;
      unsigned char* lo = hi + half_bytes;
      for (int c = ct; c < n_chunks; c += kConvWarps * 32) {
        float4 v = *reinterpret_cast<const float4*>(hi + 16 * c);
        float4 h, l;
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t); l.x = v.x - h.x;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t); l.y = v.y - h.y;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t); l.z = v.z - h.z;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t); l.w = v.w - h.w;
        *reinterpret_cast<float4*>(hi + 16 * c) = h;
        *reinterpret_cast<float4*>(lo + 16 * c) = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_conv + 8 * s);
    }
    if (nkb > 0) {
      mbar_wait(bar_acc, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = m0 + q * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* prow = p.partial + ((size_t)split * p.M + (size_t)row) * p.N;
    for (int ch = half; ch < BN / 32; ch += 2) {
      const int col0 = n0 + ch * 32;
      if (col0 >= p.N) break;
      float acc[32];
      if (nkb > 0) tmem_ld32(tlane + (uint32_t)(ch * 32), acc);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
      }
      if (row < p.M) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) prow[col0 + j] = acc[j];
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, float* C, int ldc,
                                    int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(long long)z * M * N + i];
  float* dst = C + (long long)m * ldc + n;
  *dst = accumulate ? *dst + s : s;
}

// Bt[n][k] = B[k][n]   (weights only: at most a few hundred rows/columns)
__global__ void transpose_kernel(const float* __restrict__ B, int ldb, int K, int N, float* __restrict__ Bt, int ldbt) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? B[(size_t)k * ldb + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) Bt[(size_t)n * ldbt + k] = tile[threadIdx.x][i];
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D fp32 tensor [rows][cols] (row stride ld floats), box = 32 columns x box_rows rows, 128B swizzle
bool make_map(CUtensorMap* m, const float* base, long long rows, int cols, int ld, int box_rows,
              CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void pick_bn(int N, int* bn, int* n_tiles) {
  const int t = (N + kMaxBN - 1) / kMaxBN;
  const int per = (N + t - 1) / t;
  *bn = ((per + 31) / 32) * 32;
  *n_tiles = (N + *bn - 1) / *bn;
}

size_t smem_bytes(int BN) { return (size_t)kStages * 2 * (BM * 128 + BN * 128) + (3 * kStages + 1) * 8 + 16 + 1024; }

int launch(gcnb_ctx* ctx, const TcParams& p) {
  const size_t smem = smem_bytes(p.BN);
  GCNB_CUDA(ctx, cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long m_tiles = (p.M + BM - 1) / BM;
  gemm_tc_kernel<<<(unsigned)(m_tiles * p.n_tiles), kThreads, smem, ctx->stream>>>(p);
  GCNB_LAUNCHED(ctx);
  ctx->tc_launches++;
  return GCNB_OK;
}

int transpose_into_ws(gcnb_ctx* ctx, const float* B, int ldb, int K, int N, float* Bt, int ldbt) {
  dim3 grid(cdiv(N, 32), cdiv(K, 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, ctx->stream>>>(B, ldb, K, N, Bt, ldbt);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

inline int ld32(int k) { return ((k + 31) / 32) * 32; }

}  // namespace

// ---------------------------------------------------------------------------------------------
bool gcnb_gemm_tc_supported(const gcnb_ctx* ctx, int transA, int transB, int M, int N, int K, int lda, int ldb,
                            int ldc, int accumulate) {
  (void)ctx; (void)transB; (void)accumulate; (void)ldc;
  if (transA) return false;  // wgrad (reduction over the node dimension) has its own kernel
  if (M < 1 || N < 1 || K < 1) return false;
  if ((lda % 4) != 0 || (ldb % 4) != 0) return false;
  return encode_fn() != nullptr;
}

size_t gcnb_gemm_tc_workspace_bytes(int N, int K) { return (size_t)N * ld32(K) * sizeof(float); }

int gcnb_gemm_tc(gcnb_ctx* ctx, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                 float* C, int ldc, const float* bias, int act, int accumulate) {
  GCNB_REQUIRE(ctx, aligned16(A) && aligned16(B) && aligned16(C), "tcgen05 gemm: 16-byte aligned matrices");
  GCNB_REQUIRE(ctx, (ldc % 4) == 0, "tcgen05 gemm: ldc multiple of 4");
  TcParams p;
  memset(&p, 0, sizeof(p));
  pick_bn(N, &p.BN, &p.n_tiles);
  const float* Bt = B;
  int ldbt = ldb;
  if (!transB) {  // B is K x N: the kernel wants N x K
    const size_t need = gcnb_gemm_tc_workspace_bytes(N, K);
    if (!ctx->ws || ctx->ws_bytes < need)
      return gcnb_fail(ctx, GCNB_E_WORKSPACE, "tcgen05 gemm needs %s%lld workspace bytes, have %lld", "",
                       (long long)need, (long long)ctx->ws_bytes);
    ldbt = ld32(K);
    float* w = reinterpret_cast<float*>(ctx->ws);
    int rc = transpose_into_ws(ctx, B, ldb, K, N, w, ldbt);
    if (rc != GCNB_OK) return rc;
    Bt = w;
  }
  if (!make_map(&p.mapA[0], A, M, K, lda, BM) || !make_map(&p.mapB[0], Bt, N, K, ldbt, p.BN))
    return gcnb_fail(ctx, GCNB_E_CUDA, "cuTensorMapEncodeTiled failed%s", "");
  p.nphase = 1;
  p.kblocks[0] = cdiv(K, BK);
  p.M = M; p.N = N;
  p.C = C; p.ldc = ldc; p.bias = bias; p.act = act; p.accumulate = accumulate;
  return launch(ctx, p);
}

bool gcnb_highway_tc_supported(const gcnb_ctx* ctx, int n_rows, int hd, int lds, int ldx, int ldwh, int ldwt) {
  (void)ctx;
  if (n_rows < 1 || hd < 1) return false;
  if ((lds % 4) || (ldx % 4) || (ldwh % 4) || (ldwt % 4)) return false;
  return encode_fn() != nullptr;
}

size_t gcnb_highway_tc_workspace_bytes(int hd) { return 2 * (size_t)hd * ld32(hd) * sizeof(float); }

int gcnb_highway_tc(gcnb_ctx* ctx, int n_rows, int hd, const float* S, int lds, const float* X, int ldx,
                    const float* Wh, int ldwh, const float* bh, const float* Wt, int ldwt, const float* bt, int act,
                    float* Y, int ldy, float* H, int ldh, float* T, int ldt) {
  GCNB_REQUIRE(ctx, aligned16(S) && aligned16(X) && aligned16(Y) && aligned16(Wh) && aligned16(Wt),
               "tcgen05 highway: 16-byte aligned matrices");
  const size_t need = gcnb_highway_tc_workspace_bytes(hd);
  if (!ctx->ws || ctx->ws_bytes < need)
    return gcnb_fail(ctx, GCNB_E_WORKSPACE, "tcgen05 highway needs %s%lld workspace bytes, have %lld", "",
                     (long long)need, (long long)ctx->ws_bytes);
  const int ldw = ld32(hd);
  float* WhT = reinterpret_cast<float*>(ctx->ws);
  float* WtT = WhT + (size_t)hd * ldw;
  int rc = transpose_into_ws(ctx, Wh, ldwh, hd, hd, WhT, ldw);
  if (rc != GCNB_OK) return rc;
  rc = transpose_into_ws(ctx, Wt, ldwt, hd, hd, WtT, ldw);
  if (rc != GCNB_OK) return rc;
  TcParams p;
  memset(&p, 0, sizeof(p));
  pick_bn(hd, &p.BN, &p.n_tiles);
  if (!make_map(&p.mapA[0], S, n_rows, hd, lds, BM) || !make_map(&p.mapB[0], WhT, hd, hd, ldw, p.BN) ||
      !make_map(&p.mapA[1], X, n_rows, hd, ldx, BM) || !make_map(&p.mapB[1], WtT, hd, hd, ldw, p.BN))
    return gcnb_fail(ctx, GCNB_E_CUDA, "cuTensorMapEncodeTiled failed%s", "");
  p.nphase = 2;
  p.kblocks[0] = p.kblocks[1] = cdiv(hd, BK);
  p.M = n_rows; p.N = hd;
  p.C = Y; p.ldc = ldy; p.bias = bh; p.act = act; p.accumulate = 0;
  p.bias_t = bt; p.X = X; p.ldx = ldx; p.H = H; p.ldh = ldh; p.T = T; p.ldt = ldt;
  return launch(ctx, p);
}

// ---------------------------------------------------------------------------------------------
namespace {
struct WgPlan { int BN, m_tiles, n_tiles, splits, kb_per_split; };
WgPlan wgrad_plan(int sm_count, int M, int N, int K) {
  WgPlan w;
  pick_bn(N, &w.BN, &w.n_tiles);
  w.m_tiles = (M + BM - 1) / BM;
  const int tiles = w.m_tiles * w.n_tiles;
  const int kb_total = (K + BK - 1) / BK;
  int splits = sm_count / tiles;
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total > 0 ? kb_total : 1;
  w.kb_per_split = (kb_total + splits - 1) / splits;
  if (w.kb_per_split < 1) w.kb_per_split = 1;
  w.splits = (kb_total + w.kb_per_split - 1) / w.kb_per_split;
  if (w.splits < 1) w.splits = 1;
  return w;
}
}  // namespace

bool gcnb_wgrad_tc_supported(const gcnb_ctx* ctx, int M, int N, int K, int lda, int ldb) {
  (void)ctx;
  if (M < 1 || N < 1 || K < 1) return false;
  if (M > 1024 || N > 1024) return false;  // weight-shaped outputs only
  if ((lda % 4) != 0 || (ldb % 4) != 0) return false;
  return encode_fn() != nullptr;
}

size_t gcnb_wgrad_tc_workspace_bytes(int M, int N, int K) {
  const WgPlan w = wgrad_plan(148, M, N, K);
  return (size_t)(w.splits + 1) * M * N * sizeof(float);
}

int gcnb_wgrad_tc(gcnb_ctx* ctx, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                  int ldc, int accumulate) {
  GCNB_REQUIRE(ctx, aligned16(A) && aligned16(B), "tcgen05 wgrad: 16-byte aligned matrices");
  const WgPlan w = wgrad_plan(ctx->sm_count, M, N, K);
  const size_t need = (size_t)w.splits * M * N * sizeof(float);
  if (!ctx->ws || ctx->ws_bytes < need)
    return gcnb_fail(ctx, GCNB_E_WORKSPACE, "tcgen05 wgrad needs %s%lld workspace bytes, have %lld", "",
                     (long long)need, (long long)ctx->ws_bytes);
  WgParams p;
  memset(&p, 0, sizeof(p));
  if (!make_map(&p.mapA, A, K, M, lda, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
      !make_map(&p.mapB, B, K, N, ldb, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
    return gcnb_fail(ctx, GCNB_E_CUDA, "cuTensorMapEncodeTiled failed%s", "");
  p.M = M; p.N = N; p.K = K; p.BN = w.BN; p.m_tiles = w.m_tiles; p.n_tiles = w.n_tiles;
  p.splits = w.splits; p.kb_per_split = w.kb_per_split;
  p.partial = reinterpret_cast<float*>(ctx->ws);
  const size_t smem = smem_bytes(p.BN);
  GCNB_CUDA(ctx, cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  wgrad_tc_kernel<<<(unsigned)(w.m_tiles * w.n_tiles * w.splits), kThreads, smem, ctx->stream>>>(p);
  GCNB_LAUNCHED(ctx);
  ctx->tc_launches++;
  const long long n = (long long)M * N;
  wgrad_reduce_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(p.partial, w.splits, M, N, C, ldc, accumulate);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}
