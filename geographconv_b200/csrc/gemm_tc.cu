// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMMs for sm_100a with fp32-grade accuracy.
//
// Replaces T.dot (reference gcnmodel.py:126,149,285) and its dgrad products on the hot path, and
// fuses the whole highway layer (gcnmodel.py:266,281-288) into one kernel:
//     C = epilogue(A . Bt^T)          A: M x K row-major, Bt: N x K row-major (both "K-major")
//     highway:  h = act(S.Wh + bh), t = sigmoid(X.Wt + bt), Y = t*h + (1-t)*X   (two accumulators)
//
// Precision: the reference computes in fp32 (BLAS sgemm).  kind::tf32 alone (10-bit mantissa) misses
// the 1e-3 parity budget after a few layers, so every product is done as an error-compensated
// 3xTF32 split: a = a_hi + a_lo with a_hi = a truncated to tf32, a_lo = a - a_hi (exact in fp32);
// A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, all accumulated in fp32 in TMEM (error ~2^-21).
//
// Structure (one 128 x BN output tile per CTA, 320 threads, 1 CTA / SM):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes (128B-swizzled) of raw fp32 A and Bt
//               tiles into a 3-stage shared-memory ring, completion on `full` mbarriers;
//   warps 2-9   converters: for every landed tile write the residual `lo` = a - tf32_trunc(a) into a second
//               tile with the same swizzled layout (the raw tile is the hi operand), fence.proxy.async,
//               arrive on `conv`;
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
  CUtensorMap mapB[2];    // weights, hi part (raw fp32; the tensor core drops the low 13 mantissa bits)
  CUtensorMap mapBlo[2];  // weights, residual lo part, pre-split by split_weights_kernel
  CUtensorMap mapIn;      // X tile (highway) or C tile (accumulate): boxes of 32 columns x 128 rows
  CUtensorMap mapOut[3];  // plain: C; highway: Y, H, T
  int nphase;      // k-loops: 1 = one product; 2 = two (highway: S.Wh and X.Wt; pair: A1.B1 + A2.B2)
  int hw;          // 1: fused highway epilogue (two accumulators); 0: plain epilogue (one accumulator)
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
  int blo_in_kernel;  // 1: the converter warps derive the weights' lo tile from the hi tile in shared memory (no B_lo
                      // TMA load: 36 instead of 56 KB per k-block from L2); 0: B_lo arrives pre-split through mapBlo
  long long* dbg;  // optional per-CTA phase timestamps (tools/gemm_phases.py)
  int dbg_mode;    // tools only: 1 = identity activations, 2 = no stores, 4 = no bias loads
  int prefetch;    // gemm_tc2_kernel: L2 look-ahead for the activation rows, in 128-row tiles (0 = off)
  int bar_off;     // gemm_tc2_kernel: byte offset of the barrier block (behind max(pipeline ring, epilogue staging))
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
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
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

__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Epilogue activations: ex2.approx / rcp.approx based (abs error ~1e-7, far inside the 1e-3 parity budget); the
// accurate libm forms are ~60 instructions per element and made the 8-warp epilogue latency bound.
__device__ __forceinline__ float sigmoidf_(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }
__device__ __forceinline__ float tanhf_(float z) {
  const float e = __expf(2.f * z);          // inf for large z -> 1, 0 for very negative z -> -1
  return 1.f - __fdividef(2.f, e + 1.f);
}

__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte alignment of every tile (SWIZZLE_128B atoms are 8 rows x 128 B)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const long long t_enter = clock64();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  const uint32_t a_bytes = BM * 128, b_bytes = (uint32_t)BN * 128;
  const uint32_t half_bytes = a_bytes + b_bytes;   // [A_hi | B_hi] then [A_lo | B_lo]
  const uint32_t stage_bytes = 2 * half_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * stage_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 2);
  // [2][kMaxBN] bias (bh) and gate bias (bt) of this tile; 16-byte aligned (barriers end at +88, the slot at +92)
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 6);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_conv = bar_full + 8 * kStages, bar_empty = bar_conv + 8 * kStages;
  const uint32_t bar_acc = bar_empty + 8 * kStages, bar_in = bar_acc + 8;

  const int n_tile = blockIdx.x % p.n_tiles, m_tile = blockIdx.x / p.n_tiles;
  const int m0 = m_tile * BM, n0 = n_tile * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, kConvWarps);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_in, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 64 && (int)threadIdx.x - 64 < kMaxBN) {
    // the tile's bias slices, read once per CTA: the epilogue takes them from shared memory (broadcast reads) instead
    // of 16 dependent global loads per thread and chunk (4.8k of the 22k epilogue clocks, tools/gemm_phases.py)
    const int c = (int)threadIdx.x - 64, col = n0 + c;
    const bool on = c < BN && col < p.N && !(p.dbg_mode & 4);
    const bool use_bias = p.hw || (!p.accumulate && p.bias != nullptr);
    bias_s[c] = (on && use_bias) ? __ldg(p.bias + col) : 0.f;
    bias_s[kMaxBN + c] = (on && p.hw) ? __ldg(p.bias_t + col) : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int total_kb = p.kblocks[0] + (p.nphase > 1 ? p.kblocks[1] : 0);
  long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 16 : nullptr;
  if (dbg && threadIdx.x == 64) { dbg[0] = t_enter; dbg[1] = clock64(); unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); dbg[6] = sm; dbg[7] = globaltimer_ns(); }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      long long w_empty = 0;
      for (int ph = 0; ph < p.nphase; ++ph) {
        for (int kb = 0; kb < p.kblocks[ph]; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t par = (it / kStages) & 1;
          const long long t0 = dbg ? clock64() : 0;
          mbar_wait(bar_empty + 8 * s, par ^ 1);
          if (dbg) { w_empty += clock64() - t0; dbg[8] = w_empty; }
          mbar_expect_tx(bar_full + 8 * s, a_bytes + (p.blo_in_kernel ? 1 : 2) * b_bytes);
          const uint32_t dst = smem_base + s * stage_bytes;
          tma_load_2d(dst, &p.mapA[ph], kb * BK, m0, bar_full + 8 * s);
          tma_load_2d(dst + a_bytes, &p.mapB[ph], kb * BK, n0, bar_full + 8 * s);
          if (!p.blo_in_kernel) tma_load_2d(dst + half_bytes + a_bytes, &p.mapBlo[ph], kb * BK, n0, bar_full + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc_tf32(BN);
    int it = 0;
    long long w_conv = 0;
    for (int ph = 0; ph < p.nphase; ++ph) {
      const uint32_t tacc = tmem_base + (uint32_t)(p.hw ? ph * BN : 0);  // pair products share one accumulator
      for (int kb = 0; kb < p.kblocks[ph]; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t par = (it / kStages) & 1;
        const long long t0 = dbg ? clock64() : 0;
        mbar_wait(bar_conv + 8 * s, par);
        if (dbg && lane == 0) { w_conv += clock64() - t0; dbg[9] = w_conv; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t a_hi = smem_base + s * stage_bytes, b_hi = a_hi + a_bytes;
          const uint32_t a_lo = a_hi + half_bytes, b_lo = b_hi + half_bytes;
          // small terms first, then hi.hi
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)
            umma_tf32(tacc, umma_desc_k128(a_lo + 32 * k), umma_desc_k128(b_hi + 32 * k), idesc,
                      (kb | k) != 0 || (ph != 0 && !p.hw));
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
    long long w_full = 0, t_cv = 0;
    for (int it = 0; it < total_kb; ++it) {
      const int s = it % kStages;
      const uint32_t par = (it / kStages) & 1;
      const long long t0 = dbg ? clock64() : 0;
      mbar_wait(bar_full + 8 * s, par);
      const long long t1 = dbg ? clock64() : 0;
      if (dbg && ct == 0) { w_full += t1 - t0; dbg[10] = w_full; }
      if (dbg && ct == 0 && it == 0) dbg[2] = clock64();
      // Only the activation tile needs splitting here (the weights arrive pre-split): 128 rows x 128 B = 1024
      // 16-byte chunks, 4 per thread.  The tensor core reads tf32 operands from the fp32 words as they are and
      // ignores the low 13 mantissa bits, so the raw tile IS the hi operand; only lo = a - hi is written.
      const unsigned char* hi = smem + (size_t)s * stage_bytes;
      unsigned char* lo = smem + (size_t)s * stage_bytes + half_bytes;
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(hi + 16 * (ct + 256 * u));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float4 l;
        l.x = v[u].x - __uint_as_float(__float_as_uint(v[u].x) & 0xffffe000u);
        l.y = v[u].y - __uint_as_float(__float_as_uint(v[u].y) & 0xffffe000u);
        l.z = v[u].z - __uint_as_float(__float_as_uint(v[u].z) & 0xffffe000u);
        l.w = v[u].w - __uint_as_float(__float_as_uint(v[u].w) & 0xffffe000u);
        *reinterpret_cast<float4*>(lo + 16 * (ct + 256 * u)) = l;
      }
      if (p.blo_in_kernel) {  // the weight tile the same way: BN rows x 128 B right behind the activation tile
        const int nchunks = BN * 8;
        for (int c = ct; c < nchunks; c += 256) {
          const float4 w = *reinterpret_cast<const float4*>(hi + a_bytes + 16 * c);
          float4 l;
          l.x = w.x - __uint_as_float(__float_as_uint(w.x) & 0xffffe000u);
          l.y = w.y - __uint_as_float(__float_as_uint(w.y) & 0xffffe000u);
          l.z = w.z - __uint_as_float(__float_as_uint(w.z) & 0xffffe000u);
          l.w = w.w - __uint_as_float(__float_as_uint(w.w) & 0xffffe000u);
          *reinterpret_cast<float4*>(lo + a_bytes + 16 * c) = l;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor core reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_conv + 8 * s);
      if (dbg && ct == 0) { t_cv += clock64() - t1; dbg[11] = t_cv; }
    }

    if (dbg && ct == 0) dbg[3] = clock64();
    mbar_wait(bar_acc, 0);
    if (dbg && ct == 0) dbg[4] = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: TMEM -> registers -> swizzled shared-memory boxes -> TMA stores (coalesced, clipped at M / N).
    // The pipeline stages are free now; they are re-used as [IN boxes | OUT staging of group 0 | group 1].
    // Warps 2-5 form group 0, warps 6-9 group 1; a group owns every second 32-column chunk of the tile.
    constexpr uint32_t kBox = BM * 128;  // one box: 128 rows x 32 columns fp32, SWIZZLE_128B
    const int nch = BN / 32;
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;       // row of the tile this thread owns
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool need_in = p.hw || p.accumulate;
    const int n_out = p.hw ? 3 : 1;
    unsigned char* in_base = smem;
    unsigned char* out_base = smem + (need_in ? (size_t)nch * kBox : 0) + (size_t)grp * n_out * kBox;
    if (need_in) {
      if (ct == 0) {
        mbar_expect_tx(bar_in, (uint32_t)nch * kBox);
        for (int c = 0; c < nch; ++c) tma_load_2d(smem_base + c * kBox, &p.mapIn, n0 + 32 * c, m0, bar_in);
      }
      mbar_wait(bar_in, 0);
    }
    long long e_ld = 0, e_math = 0, e_out = 0;
    if (dbg && ct == 0) dbg[12] = clock64() - dbg[4];
    const uint32_t swz = (uint32_t)(r & 7);
    const uint32_t bar_id = 1 + grp;
    const bool leader = (ct & 127) == 0;
    bool first = true;
    for (int ch = grp; ch < nch; ch += 2) {
      const int col0 = n0 + ch * 32;
      if (col0 >= p.N) break;  // uniform over the group
      float acc[32];
      const long long s0 = dbg ? clock64() : 0;
      tmem_ld32(tlane + (uint32_t)(ch * 32), acc);
      const long long s1 = dbg ? clock64() : 0;
      float xin[32];
      if (need_in) {
        const unsigned char* ib = in_base + (size_t)ch * kBox + (size_t)r * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(ib + ((j ^ swz) << 4));
          xin[4 * j] = v.x; xin[4 * j + 1] = v.y; xin[4 * j + 2] = v.z; xin[4 * j + 3] = v.w;
        }
      }
      if (!first) {  // the previous chunk's TMA stores must have read the staging boxes
        if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      }
      first = false;
      unsigned char* ob = out_base + (size_t)r * 128;
      const float4* bs = reinterpret_cast<const float4*>(bias_s + ch * 32);
      if (!p.hw) {
        auto body = [&](auto act_fn) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = bs[j];
            const float b[4] = {bv.x, bv.y, bv.z, bv.w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[e] = p.accumulate ? acc[4 * j + e] + xin[4 * j + e] : act_fn(acc[4 * j + e] + b[e]);
            *reinterpret_cast<float4*>(ob + ((j ^ swz) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
          }
        };
        if (p.accumulate || p.act == GCNB_ACT_LINEAR) body([](float z) { return z; });
        else if (p.act == GCNB_ACT_TANH) body([](float z) { return tanhf_(z); });
        else if (p.act == GCNB_ACT_SIGMOID) body([](float z) { return sigmoidf_(z); });
        else body([](float z) { return fmaxf(z, 0.f); });
      } else {
        float acc_t[32];
        tmem_ld32(tlane + (uint32_t)(BN + ch * 32), acc_t);
        const float4* bts = reinterpret_cast<const float4*>(bias_s + kMaxBN + ch * 32);
        auto body = [&](auto act_fn) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 u = bs[j], w = bts[j];
            const float bh[4] = {u.x, u.y, u.z, u.w}, bt[4] = {w.x, w.y, w.z, w.w};
            float h[4], t[4], y[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              h[e] = act_fn(acc[4 * j + e] + bh[e]);
              t[e] = (p.dbg_mode & 1) ? acc_t[4 * j + e] + bt[e] : sigmoidf_(acc_t[4 * j + e] + bt[e]);
              y[e] = t[e] * h[e] + (1.0f - t[e]) * xin[4 * j + e];
            }
            const uint32_t off = (j ^ swz) << 4;
            *reinterpret_cast<float4*>(ob + off) = make_float4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<float4*>(ob + kBox + off) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(ob + 2 * kBox + off) = make_float4(t[0], t[1], t[2], t[3]);
          }
        };
        if (p.dbg_mode & 1) body([](float z) { return z; });
        else if (p.act == GCNB_ACT_TANH) body([](float z) { return tanhf_(z); });
        else if (p.act == GCNB_ACT_SIGMOID) body([](float z) { return sigmoidf_(z); });
        else if (p.act == GCNB_ACT_RELU) body([](float z) { return fmaxf(z, 0.f); });
        else body([](float z) { return z; });
      }
      const long long s2 = dbg ? clock64() : 0;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (dbg && ct == 0) { e_ld += s1 - s0; e_math += s2 - s1; e_out += clock64() - s2; dbg[13] = e_ld; dbg[14] = e_math; dbg[15] = e_out; }
      if (leader && !(p.dbg_mode & 2)) {
        const uint32_t src = smem_u32(out_base);
        tma_store_2d(&p.mapOut[0], src, col0, m0);
        if (p.hw) {
          if (p.H) tma_store_2d(&p.mapOut[1], src + kBox, col0, m0);
          if (p.T) tma_store_2d(&p.mapOut[2], src + 2 * kBox, col0, m0);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (dbg && ct == 0) dbg[5] = clock64();
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// gemm_tc2_kernel -- the same products with TWO co-resident CTAs per SM (default, ctx->gemm_v = 2).
//
// Measured on the kernel above (profiles/r1c_gemm_phases.txt): a highway tile spends 26.7k clocks in the main loop
// (3xTF32 MMA floor 19.2k) and then 18.9k in an epilogue during which the tensor pipe idles and HBM carries the whole
// output traffic -- 37% tensor-pipe activity.  Instead of a persistent kernel with dedicated epilogue warps this kernel
// halves every per-CTA resource so that two CTAs share an SM and the hardware overlaps one CTA's epilogue (HBM, MUFU)
// with the other's main loop (L2 -> shared memory, tensor pipe):
//   * 256 TMEM columns per CTA: highway tiles are 128 columns wide (two accumulators), plain tiles up to 160; the last
//     n-tile of a row block only issues MMAs for its own width rounded up to 16 (N = 300: 128 + 128 + 48 columns);
//   * k-blocks of 16 floats (64-byte rows, SWIZZLE_64B boxes and descriptors): a stage is 2 x (8 KB + BN x 64 B), three
//     stages fit 96-108 KB, 6 tcgen05.mma per k-block;
//   * <= 96 registers per thread: the epilogue works on 16 columns per thread (two warps share a TMEM lane quarter and
//     split every 32-column chunk), the X / C tile arrives chunk by chunk in two alternating boxes, outputs leave
//     through one set of 32-column staging boxes and TMA stores.
// ---------------------------------------------------------------------------------------------
constexpr int BK2 = 16;
constexpr int kStages2 = 3;
constexpr int kTmemCols2 = 256;
constexpr int kMaxBN2 = 160;     // one accumulator
constexpr int kMaxBN2hw = 128;   // two accumulators in 256 columns
constexpr int kBarBlock2 = 112;  // 12 barriers (96 B), TMEM slot, padding to 16

// K-major SWIZZLE_64B descriptor: rows of 64 B, 8-row groups 512 B apart (SBO), layout type 4
__device__ __forceinline__ uint64_t umma_desc_k64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(kThreads, 2) gemm_tc2_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const long long t_enter = clock64();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;  // tile pitch = rows of a weight box; this tile's own MMA width is bn
  const int n_tile = blockIdx.x % p.n_tiles, m_tile = blockIdx.x / p.n_tiles;
  const int m0 = m_tile * BM, n0 = n_tile * BN;
  int bn = p.N - n0;
  bn = bn >= BN ? BN : ((bn + 15) & ~15);
  const uint32_t a_bytes = BM * 64, b_bytes = (uint32_t)BN * 64;
  const uint32_t half_bytes = a_bytes + b_bytes;  // [A_hi | B_hi] then [A_lo | B_lo]
  const uint32_t stage_bytes = 2 * half_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* bias_s = reinterpret_cast<float*>(smem + p.bar_off + kBarBlock2);  // [2][kMaxBN2]
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_conv = bar_full + 8 * kStages2, bar_empty = bar_conv + 8 * kStages2;
  const uint32_t bar_acc = bar_empty + 8 * kStages2, bar_in = bar_acc + 8;  // bar_in, bar_in + 8: the two IN boxes

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, kConvWarps);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_in, 1);
    mbar_init(bar_in + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols2)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 64 && (int)threadIdx.x - 64 < kMaxBN2) {
    const int c = (int)threadIdx.x - 64, col = n0 + c;
    const bool on = c < bn && col < p.N && !(p.dbg_mode & 4);
    const bool use_bias = p.hw || (!p.accumulate && p.bias != nullptr);
    bias_s[c] = (on && use_bias) ? __ldg(p.bias + col) : 0.f;
    bias_s[kMaxBN2 + c] = (on && p.hw) ? __ldg(p.bias_t + col) : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int total_kb = p.kblocks[0] + (p.nphase > 1 ? p.kblocks[1] : 0);
  long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 16 : nullptr;
  if (dbg && threadIdx.x == 64) { dbg[0] = t_enter; dbg[1] = clock64(); unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); dbg[6] = sm; dbg[7] = globaltimer_ns(); }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // The activation rows stream from HBM exactly once (the other n-tiles of a row block hit L2).  With three stages
      // per CTA a DRAM round trip per k-block would stall the ring, so the CTA of n-tile 0 asks L2 for the rows of the
      // row block `prefetch` tiles ahead (about what the grid has in flight): by the time those CTAs start, their
      // operands are L2 hits.
      if (p.prefetch > 0 && n_tile == 0) {
        const long long m_ahead = (long long)m0 + (long long)p.prefetch * BM;
        if (m_ahead < p.M)
          for (int ph = 0; ph < p.nphase; ++ph)
            for (int kb = 0; kb < p.kblocks[ph]; ++kb) tma_prefetch_2d(&p.mapA[ph], kb * BK2, (int)m_ahead);
      }
      int it = 0;
      long long w_empty = 0;
      for (int ph = 0; ph < p.nphase; ++ph) {
        for (int kb = 0; kb < p.kblocks[ph]; ++kb, ++it) {
          const int s = it % kStages2;
          const uint32_t par = (it / kStages2) & 1;
          const long long t0 = dbg ? clock64() : 0;
          mbar_wait(bar_empty + 8 * s, par ^ 1);
          if (dbg) { w_empty += clock64() - t0; dbg[8] = w_empty; }
          mbar_expect_tx(bar_full + 8 * s, a_bytes + (p.blo_in_kernel ? 1 : 2) * b_bytes);
          const uint32_t dst = smem_base + s * stage_bytes;
          tma_load_2d(dst, &p.mapA[ph], kb * BK2, m0, bar_full + 8 * s);
          tma_load_2d(dst + a_bytes, &p.mapB[ph], kb * BK2, n0, bar_full + 8 * s);
          if (!p.blo_in_kernel) tma_load_2d(dst + half_bytes + a_bytes, &p.mapBlo[ph], kb * BK2, n0, bar_full + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc_tf32(bn);
    int it = 0;
    long long w_conv = 0;
    for (int ph = 0; ph < p.nphase; ++ph) {
      const uint32_t tacc = tmem_base + (uint32_t)(p.hw ? ph * BN : 0);  // pair products share one accumulator
      for (int kb = 0; kb < p.kblocks[ph]; ++kb, ++it) {
        const int s = it % kStages2;
        const uint32_t par = (it / kStages2) & 1;
        // Terms that read raw tiles only (the raw tile is the hi operand) are issued as soon as the TMA bytes land and
        // run while the converter warps write the residual tiles; the terms with a residual operand follow.
        // blo_in_kernel: both residuals are made here (A_lo and B_lo), so hi.hi goes first and the two cross terms wait;
        // otherwise B_lo arrived by TMA and only A_lo.B_hi waits.
        mbar_wait(bar_full + 8 * s, par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_base + s * stage_bytes, b_hi = a_hi + a_bytes;
        const uint32_t a_lo = a_hi + half_bytes, b_lo = b_hi + half_bytes;
        const uint32_t first = (kb != 0 || (ph != 0 && !p.hw)) ? 1u : 0u;  // 0: the tile's very first MMA overwrites
        if (lane == 0) {
          if (!p.blo_in_kernel) {
#pragma unroll
            for (int k = 0; k < BK2 / 8; ++k)
              umma_tf32(tacc, umma_desc_k64(a_hi + 32 * k), umma_desc_k64(b_lo + 32 * k), idesc, k ? 1u : first);
          }
#pragma unroll
          for (int k = 0; k < BK2 / 8; ++k)
            umma_tf32(tacc, umma_desc_k64(a_hi + 32 * k), umma_desc_k64(b_hi + 32 * k), idesc,
                      (k || !p.blo_in_kernel) ? 1u : first);
        }
        __syncwarp();
        const long long t0 = dbg ? clock64() : 0;
        mbar_wait(bar_conv + 8 * s, par);
        if (dbg && lane == 0) { w_conv += clock64() - t0; dbg[9] = w_conv; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < BK2 / 8; ++k)
            umma_tf32(tacc, umma_desc_k64(a_lo + 32 * k), umma_desc_k64(b_hi + 32 * k), idesc, 1u);
          if (p.blo_in_kernel) {
#pragma unroll
            for (int k = 0; k < BK2 / 8; ++k)
              umma_tf32(tacc, umma_desc_k64(a_hi + 32 * k), umma_desc_k64(b_lo + 32 * k), idesc, 1u);
          }
          umma_commit(bar_empty + 8 * s);  // implies tcgen05.fence::before_thread_sync
          if (it == total_kb - 1) umma_commit(bar_acc);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ converters, then epilogue
    const int ct = threadIdx.x - 64;  // 0 .. 255
    long long w_full = 0, t_cv = 0;
    for (int it = 0; it < total_kb; ++it) {
      const int s = it % kStages2;
      const uint32_t par = (it / kStages2) & 1;
      const long long t0 = dbg ? clock64() : 0;
      mbar_wait(bar_full + 8 * s, par);
      const long long t1 = dbg ? clock64() : 0;
      if (dbg && ct == 0) { w_full += t1 - t0; dbg[10] = w_full; }
      if (dbg && ct == 0 && it == 0) dbg[2] = clock64();
      // activation tile: 128 rows x 64 B = 512 16-byte chunks, 2 per thread; the raw tile is the hi operand (the tensor
      // core ignores the low 13 mantissa bits), only lo = a - hi is written, in the same swizzled positions.
      // blo_in_kernel: the weight tile right behind it (bn rows x 64 B) the same way -- its residual then never crosses
      // the L2 -> SM fabric, which is what bounds the main loop (43 B / clk / SM measured in both kernels).
      const unsigned char* hi = smem + (size_t)s * stage_bytes;
      unsigned char* lo = smem + (size_t)s * stage_bytes + half_bytes;
      const int n_conv = 512 + (p.blo_in_kernel ? bn * 4 : 0);
      for (int c0 = 0; c0 < n_conv; c0 += 512) {
        float4 v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
          v[u] = (c0 + ct + 256 * u < n_conv) ? *reinterpret_cast<const float4*>(hi + 16 * (c0 + ct + 256 * u))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float4 l;
          l.x = v[u].x - __uint_as_float(__float_as_uint(v[u].x) & 0xffffe000u);
          l.y = v[u].y - __uint_as_float(__float_as_uint(v[u].y) & 0xffffe000u);
          l.z = v[u].z - __uint_as_float(__float_as_uint(v[u].z) & 0xffffe000u);
          l.w = v[u].w - __uint_as_float(__float_as_uint(v[u].w) & 0xffffe000u);
          if (c0 + ct + 256 * u < n_conv) *reinterpret_cast<float4*>(lo + 16 * (c0 + ct + 256 * u)) = l;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor core reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_conv + 8 * s);
      if (dbg && ct == 0) { t_cv += clock64() - t1; dbg[11] = t_cv; }
    }

    if (dbg && ct == 0) dbg[3] = clock64();
    mbar_wait(bar_acc, 0);
    if (dbg && ct == 0) dbg[4] = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: TMEM -> registers -> swizzled staging boxes -> TMA stores (coalesced, clipped at M / N).
    // The ring is free now: [IN box 0 | IN box 1 | OUT boxes].  All eight warps work on the same 32-column chunk: warps
    // 2-5 on its first 16 columns, warps 6-9 on the last 16 (two warps share each TMEM lane quarter).
    constexpr uint32_t kBox = BM * 128;  // 128 rows x 32 columns fp32, SWIZZLE_128B
    const int nch = (bn + 31) / 32;
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;   // which half of a chunk
    const int r = q * 32 + lane;       // row of the tile this thread owns
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * 16);
    const bool need_in = p.hw || p.accumulate;
    unsigned char* out_base = smem + (need_in ? 2 * (size_t)kBox : 0);
    int n_valid = 0;  // chunks that start inside the matrix
    for (int c = 0; c < nch; ++c) n_valid += (n0 + 32 * c < p.N);
    if (need_in && ct == 0) {
      for (int c = 0; c < 2 && c < n_valid; ++c) {
        mbar_expect_tx(bar_in + 8 * c, kBox);
        tma_load_2d(smem_base + c * kBox, &p.mapIn, n0 + 32 * c, m0, bar_in + 8 * c);
      }
    }
    long long e_in = 0, e_math = 0, e_out = 0;
    const uint32_t swz = (uint32_t)(r & 7);
    for (int ch = 0; ch < n_valid; ++ch) {
      const int col0 = n0 + ch * 32;
      float acc[16];
      tmem_ld16(tlane + (uint32_t)(ch * 32), acc);
      const long long s0 = dbg ? clock64() : 0;
      float xin[16];
      if (need_in) {
        mbar_wait(bar_in + 8 * (ch & 1), (uint32_t)(ch >> 1) & 1);
        const unsigned char* ib = smem + (size_t)(ch & 1) * kBox + (size_t)r * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(ib + (((uint32_t)(4 * grp + j) ^ swz) << 4));
          xin[4 * j] = v.x; xin[4 * j + 1] = v.y; xin[4 * j + 2] = v.z; xin[4 * j + 3] = v.w;
        }
      }
      const long long s1 = dbg ? clock64() : 0;
      // all arithmetic happens in place in registers BEFORE the wait for the staging boxes, so the MUFU work of this
      // chunk overlaps the TMA stores of the previous one: acc -> output (plain) or h, acc_t -> t, xin -> y (highway)
      const float4* bs = reinterpret_cast<const float4*>(bias_s + ch * 32 + grp * 16);
      float acc_t[16];
      if (!p.hw) {
        auto body = [&](auto act_fn) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 bv = bs[j];
            const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              acc[4 * j + e] = p.accumulate ? acc[4 * j + e] + xin[4 * j + e] : act_fn(acc[4 * j + e] + b[e]);
          }
        };
        if (p.accumulate || p.act == GCNB_ACT_LINEAR) body([](float z) { return z; });
        else if (p.act == GCNB_ACT_TANH) body([](float z) { return tanhf_(z); });
        else if (p.act == GCNB_ACT_SIGMOID) body([](float z) { return sigmoidf_(z); });
        else body([](float z) { return fmaxf(z, 0.f); });
      } else {
        tmem_ld16(tlane + (uint32_t)(BN + ch * 32), acc_t);
        const float4* bts = reinterpret_cast<const float4*>(bias_s + kMaxBN2 + ch * 32 + grp * 16);
        auto body = [&](auto pair_fn) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 u = bs[j], w = bts[j];
            const float bh[4] = {u.x, u.y, u.z, u.w}, bt[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float h, t;
              pair_fn(acc[4 * j + e] + bh[e], acc_t[4 * j + e] + bt[e], h, t);
              acc[4 * j + e] = h;
              acc_t[4 * j + e] = t;
              xin[4 * j + e] = t * h + (1.0f - t) * xin[4 * j + e];
            }
          }
        };
        if (p.dbg_mode & 1) body([](float zh, float zt, float& h, float& t) { h = zh; t = zt; });
        else if (p.act == GCNB_ACT_TANH)
          // tanh(zh) = 1 - 2/(e1 + 1), sigmoid(zt) = 1/(e2 + 1) with e1 = exp(2 zh), e2 = exp(-zt): ONE reciprocal serves
          // both, r = 1/((e1 + 1)(e2 + 1)).  Exponents are clamped at 40 (tanh is 1 to fp32 precision beyond 2 zh = 18,
          // the gate below 4e-18) so the product stays finite; 3 MUFU operations per element instead of 4.
          body([](float zh, float zt, float& h, float& t) {
            const float a = __expf(fminf(2.f * zh, 40.f)) + 1.f, b = __expf(fminf(-zt, 40.f)) + 1.f;
            const float r = __fdividef(1.f, a * b);
            h = 1.f - 2.f * (r * b);
            t = r * a;
          });
        else if (p.act == GCNB_ACT_SIGMOID) body([](float zh, float zt, float& h, float& t) { h = sigmoidf_(zh); t = sigmoidf_(zt); });
        else if (p.act == GCNB_ACT_RELU) body([](float zh, float zt, float& h, float& t) { h = fmaxf(zh, 0.f); t = sigmoidf_(zt); });
        else body([](float zh, float zt, float& h, float& t) { h = zh; t = sigmoidf_(zt); });
      }
      if (ch > 0) {  // the previous chunk's TMA stores must have read the staging boxes
        if (ct == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      unsigned char* ob = out_base + (size_t)r * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = ((uint32_t)(4 * grp + j) ^ swz) << 4;
        if (!p.hw) {
          *reinterpret_cast<float4*>(ob + off) = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        } else {
          *reinterpret_cast<float4*>(ob + off) = make_float4(xin[4 * j], xin[4 * j + 1], xin[4 * j + 2], xin[4 * j + 3]);
          *reinterpret_cast<float4*>(ob + kBox + off) = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          *reinterpret_cast<float4*>(ob + 2 * kBox + off) =
              make_float4(acc_t[4 * j], acc_t[4 * j + 1], acc_t[4 * j + 2], acc_t[4 * j + 3]);
        }
      }
      const long long s2 = dbg ? clock64() : 0;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");  // staging complete; every thread has read IN box ch & 1
      if (ct == 0) {
        if (!(p.dbg_mode & 2)) {
          const uint32_t src = smem_u32(out_base);
          tma_store_2d(&p.mapOut[0], src, col0, m0);
          if (p.hw) {
            if (p.H) tma_store_2d(&p.mapOut[1], src + kBox, col0, m0);
            if (p.T) tma_store_2d(&p.mapOut[2], src + 2 * kBox, col0, m0);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (need_in && ch + 2 < n_valid) {
          mbar_expect_tx(bar_in + 8 * (ch & 1), kBox);
          tma_load_2d(smem_base + (ch & 1) * kBox, &p.mapIn, n0 + 32 * (ch + 2), m0, bar_in + 8 * (ch & 1));
        }
      }
      if (dbg && ct == 0) { e_in += s1 - s0; e_math += s2 - s1; e_out += clock64() - s2; dbg[13] = e_in; dbg[14] = e_math; dbg[15] = e_out; }
    }
    if (ct == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (dbg && ct == 0) dbg[5] = clock64();
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols2) : "memory");
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

// BKW = k-rows per pipeline stage, TMEMC = TMEM columns, MINB = CTAs per SM: <32, 512, 1> is the one-CTA-per-SM kernel,
// <16, 256, 2> (gemm_v 2) halves the stage so that two CTAs share an SM and twice as many stages are in flight.
template <int BKW, int TMEMC, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  constexpr uint32_t chunk_bytes = BKW * 128;  // one 32-column chunk: BKW k-rows x 128 B
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
  const int kb_total = (p.K + BKW - 1) / BKW;
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
                 "n"(TMEMC)
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
        const int k0 = (kb0 + it) * BKW;
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
      // the hi.hi term reads the raw tiles only: issued when the TMA bytes land, it runs while the converters write the
      // two residual tiles; the two cross terms follow
      const uint32_t a_hi = smem_base + s * stage_bytes, b_hi = a_hi + a_bytes;
      const uint32_t a_lo = a_hi + half_bytes, b_lo = b_hi + half_bytes;
      mbar_wait(bar_full + 8 * s, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < BKW / 8; ++k)
          umma_tf32(tmem_base, umma_desc_mn128(a_hi + 1024 * k, chunk_bytes), umma_desc_mn128(b_hi + 1024 * k, chunk_bytes), idesc, (it | k) != 0);
      }
      __syncwarp();
      mbar_wait(bar_conv + 8 * s, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < BKW / 8; ++k)
          umma_tf32(tmem_base, umma_desc_mn128(a_lo + 1024 * k, chunk_bytes), umma_desc_mn128(b_hi + 1024 * k, chunk_bytes), idesc, 1u);
#pragma unroll
        for (int k = 0; k < BKW / 8; ++k)
          umma_tf32(tmem_base, umma_desc_mn128(a_hi + 1024 * k, chunk_bytes), umma_desc_mn128(b_lo + 1024 * k, chunk_bytes), idesc, 1u);
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
        // The tensor core reads tf32 operands from the fp32 words as they are and ignores the low 13 mantissa
        // bits, so the raw tile IS the hi operand (hi = a with those bits dropped); only lo = a - hi is written.
        const float4 v = *reinterpret_cast<const float4*>(hi + 16 * c);
        float4 l;
        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEMC) : "memory");
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

// Weight operand prep (weights only: at most a few hundred rows / columns): Bt[n][k] = op(B) as an N x K
// row-major matrix (transposing when B is stored K x N) and its 3xTF32 residual Blo = Bt - tf32_trunc(Bt).
__global__ void split_weights_kernel(const float* __restrict__ B, int ldb, int K, int N, int transpose,
                                     float* __restrict__ Bt, float* __restrict__ Blo, int ldbt) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    float v = 0.f;
    if (transpose) {  // B is K x N
      const int k = k0 + i, n = n0 + threadIdx.x;
      if (k < K && n < N) v = B[(size_t)k * ldb + n];
      tile[i][threadIdx.x] = v;
    } else {          // B is N x K already
      const int n = n0 + i, k = k0 + threadIdx.x;
      if (k < K && n < N) v = B[(size_t)n * ldb + k];
      tile[threadIdx.x][i] = v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) {
      const float v = tile[threadIdx.x][i];
      Bt[(size_t)n * ldbt + k] = v;
      Blo[(size_t)n * ldbt + k] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    }
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
              CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int box_cols = BK) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
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

size_t smem_bytes(int BN) {
  return (size_t)kStages * 2 * (BM * 128 + BN * 128) + (3 * kStages + 2) * 8 + 32 + 2 * kMaxBN * sizeof(float) + 1024;
}

// Tiling of the output columns and of the k dimension for the kernel version in use (ctx->gemm_v).
struct Tiling {
  int v2, bk;
  CUtensorMapSwizzle swz;
};
Tiling pick_tiling(const gcnb_ctx* ctx, int N, int hw, TcParams* p) {
  Tiling t;
  t.v2 = ctx->gemm_v >= 2;
  if (!t.v2) {
    pick_bn(N, &p->BN, &p->n_tiles);
    t.bk = BK;
    t.swz = CU_TENSOR_MAP_SWIZZLE_128B;
    return t;
  }
  const int maxbn = hw ? kMaxBN2hw : kMaxBN2;
  if (N <= maxbn) {
    p->BN = ((N + 15) / 16) * 16;
  } else {
    const int tiles = (N + maxbn - 1) / maxbn;
    int per = (((N + tiles - 1) / tiles + 31) / 32) * 32;
    p->BN = per > maxbn ? maxbn : per;
  }
  p->n_tiles = (N + p->BN - 1) / p->BN;
  t.bk = BK2;
  t.swz = CU_TENSOR_MAP_SWIZZLE_64B;
  return t;
}
// operand tile maps: boxes of one k-block x box_rows rows
bool make_opmap(const Tiling& t, CUtensorMap* m, const float* base, long long rows, int cols, int ld, int box_rows) {
  return make_map(m, base, rows, cols, ld, box_rows, t.swz, t.bk);
}

int launch(gcnb_ctx* ctx, TcParams& p) {
  const long long m_tiles = (p.M + BM - 1) / BM;
  ctx->tc_launches++;
  if (ctx->gemm_v >= 2) {
    const size_t ring = (size_t)kStages2 * 2 * (BM * 64 + p.BN * 64);
    const size_t epi = (size_t)((p.hw || p.accumulate ? 2 : 0) + (p.hw ? 3 : 1)) * BM * 128;
    p.bar_off = (int)(ring > epi ? ring : epi);
    // look-ahead: the row blocks the grid has in flight (two CTAs per SM, n_tiles CTAs per row block)
    p.prefetch = ctx->gemm_prefetch < 0 ? (2 * ctx->sm_count + p.n_tiles - 1) / p.n_tiles : ctx->gemm_prefetch;
    const size_t smem = 1024 + (size_t)p.bar_off + kBarBlock2 + 2 * kMaxBN2 * sizeof(float);
    GCNB_CUDA(ctx, cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tc2_kernel<<<(unsigned)(m_tiles * p.n_tiles), kThreads, smem, ctx->stream>>>(p);
    GCNB_LAUNCHED(ctx);
    return GCNB_OK;
  }
  const size_t smem = smem_bytes(p.BN);
  GCNB_CUDA(ctx, cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gemm_tc_kernel<<<(unsigned)(m_tiles * p.n_tiles), kThreads, smem, ctx->stream>>>(p);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

int split_weights(gcnb_ctx* ctx, const float* B, int ldb, int K, int N, int transpose, float* Bt, float* Blo, int ldbt) {
  dim3 grid(cdiv(N, 32), cdiv(K, 32)), block(32, 8);
  split_weights_kernel<<<grid, block, 0, ctx->stream>>>(B, ldb, K, N, transpose, Bt, Blo, ldbt);
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

size_t gcnb_gemm_tc_workspace_bytes(int N, int K) { return 2 * (size_t)N * ld32(K) * sizeof(float); }

int gcnb_gemm_tc(gcnb_ctx* ctx, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                 float* C, int ldc, const float* bias, int act, int accumulate) {
  GCNB_REQUIRE(ctx, aligned16(A) && aligned16(B) && aligned16(C), "tcgen05 gemm: 16-byte aligned matrices");
  GCNB_REQUIRE(ctx, (ldc % 4) == 0, "tcgen05 gemm: ldc multiple of 4");
  TcParams p;
  memset(&p, 0, sizeof(p));
  const Tiling tl = pick_tiling(ctx, N, 0, &p);
  // weights: N x K copy (transposed when B is stored K x N) plus its residual, both in the workspace
  const size_t need = gcnb_gemm_tc_workspace_bytes(N, K);
  if (!ctx->ws || ctx->ws_bytes < need)
    return gcnb_fail(ctx, GCNB_E_WORKSPACE, "tcgen05 gemm needs %s%lld workspace bytes, have %lld", "",
                     (long long)need, (long long)ctx->ws_bytes);
  const int ldbt = ld32(K);
  float* Bt = reinterpret_cast<float*>(ctx->ws);
  float* Blo = Bt + (size_t)N * ldbt;
  int rc = split_weights(ctx, B, ldb, K, N, transB ? 0 : 1, Bt, Blo, ldbt);
  if (rc != GCNB_OK) return rc;
  if (!make_opmap(tl, &p.mapA[0], A, M, K, lda, BM) || !make_opmap(tl, &p.mapB[0], Bt, N, K, ldbt, p.BN) ||
      !make_opmap(tl, &p.mapBlo[0], Blo, N, K, ldbt, p.BN))
    return gcnb_fail(ctx, GCNB_E_CUDA, "cuTensorMapEncodeTiled failed%s", "");
  if (!make_map(&p.mapOut[0], C, M, N, ldc, BM) || (accumulate && !make_map(&p.mapIn, C, M, N, ldc, BM)))
    return gcnb_fail(ctx, GCNB_E_CUDA, "cuTensorMapEncodeTiled failed%s", "");
  p.nphase = 1;
  p.kblocks[0] = cdiv(K, tl.bk);
  p.M = M; p.N = N;
  p.C = C; p.ldc = ldc; p.bias = bias; p.act = act; p.accumulate = accumulate;
  p.blo_in_kernel = ctx->gemm_v >= 2 ? ctx->gemm_blo2 : ctx->gemm_blo;
  return launch(ctx, p);
}

// C (+)= A1.op(B1) + A2.op(B2) in one pass over C: two k-loops into one TMEM accumulator (the two dgrad products of
// a highway layer, dx += dTpre.Wt^T + V.Wh^T: one read and one write of dx instead of two)
int gcnb_gemm_pair_tc(gcnb_ctx* ctx, int transB, int M, int N, int K, const float* A1, int lda1, const float* B1, int ldb1,
                      const float* A2, int lda2, const float* B2, int ldb2, float* C, int ldc, int accumulate) {
  GCNB_REQUIRE(ctx, aligned16(A1) && aligned16(B1) && aligned16(A2) && aligned16(B2) && aligned16(C),
               "tcgen05 gemm: 16-byte aligned matrices");
  GCNB_REQUIRE(ctx, (ldc % 4) == 0, "tcgen05 gemm: ldc multiple of 4");
  TcParams p;
  memset(&p, 0, sizeof(p));
  const Tiling tl = pick_tiling(ctx, N, 0, &p);
  const size_t need = 2 * gcnb_gemm_tc_workspace_bytes(N, K);
  if (!ctx->ws || ctx->ws_bytes < need)
    return gcnb_fail(ctx, GCNB_E_WORKSPACE, "tcgen05 gemm pair needs %s%lld workspace bytes, have %lld", "",
                     (long long)need, (long long)ctx->ws_bytes);
  const int ldbt = ld32(K);
  const size_t wsz = (size_t)N * ldbt;
  float* Bt1 = reinterpret_cast<float*>(ctx->ws);
  float* Blo1 = Bt1 + wsz;
  float* Bt2 = Blo1 + wsz;
  float* Blo2 = Bt2 + wsz;
  int rc = split_weights(ctx, B1, ldb1, K, N, transB ? 0 : 1, Bt1, Blo1, ldbt);
  if (rc != GCNB_OK) return rc;
  rc = split_weights(ctx, B2, ldb2, K, N, transB ? 0 : 1, Bt2, Blo2, ldbt);
  if (rc != GCNB_OK) return rc;
  if (!make_opmap(tl, &p.mapA[0], A1, M, K, lda1, BM) || !make_opmap(tl, &p.mapB[0], Bt1, N, K, ldbt, p.BN) ||
      !make_opmap(tl, &p.mapBlo[0], Blo1, N, K, ldbt, p.BN) || !make_opmap(tl, &p.mapA[1], A2, M, K, lda2, BM) ||
      !make_opmap(tl, &p.mapB[1], Bt2, N, K, ldbt, p.BN) || !make_opmap(tl, &p.mapBlo[1], Blo2, N, K, ldbt, p.BN))
    return gcnb_fail(ctx, GCNB_E_CUDA, "cuTensorMapEncodeTiled failed%s", "");
  if (!make_map(&p.mapOut[0], C, M, N, ldc, BM) || (accumulate && !make_map(&p.mapIn, C, M, N, ldc, BM)))
    return gcnb_fail(ctx, GCNB_E_CUDA, "cuTensorMapEncodeTiled failed%s", "");
  p.nphase = 2;
  p.hw = 0;
  p.kblocks[0] = p.kblocks[1] = cdiv(K, tl.bk);
  p.M = M; p.N = N;
  p.C = C; p.ldc = ldc; p.bias = nullptr; p.act = GCNB_ACT_LINEAR; p.accumulate = accumulate;
  p.blo_in_kernel = ctx->gemm_v >= 2 ? ctx->gemm_blo2 : ctx->gemm_blo;
  return launch(ctx, p);
}

bool gcnb_highway_tc_supported(const gcnb_ctx* ctx, int n_rows, int hd, int lds, int ldx, int ldwh, int ldwt) {
  (void)ctx;
  if (n_rows < 1 || hd < 1) return false;
  if ((lds % 4) || (ldx % 4) || (ldwh % 4) || (ldwt % 4)) return false;
  return encode_fn() != nullptr;
}

size_t gcnb_highway_tc_workspace_bytes(int hd) { return 4 * (size_t)hd * ld32(hd) * sizeof(float); }

int gcnb_highway_tc(gcnb_ctx* ctx, int n_rows, int hd, const float* S, int lds, const float* X, int ldx,
                    const float* Wh, int ldwh, const float* bh, const float* Wt, int ldwt, const float* bt, int act,
                    float* Y, int ldy, float* H, int ldh, float* T, int ldt) {
  GCNB_REQUIRE(ctx, aligned16(S) && aligned16(X) && aligned16(Y) && aligned16(Wh) && aligned16(Wt) &&
                        (!H || aligned16(H)) && (!T || aligned16(T)) && (ldy % 4) == 0 && (!H || (ldh % 4) == 0) &&
                        (!T || (ldt % 4) == 0),
               "tcgen05 highway: 16-byte aligned matrices, leading dimensions multiple of 4");
  const size_t need = gcnb_highway_tc_workspace_bytes(hd);
  if (!ctx->ws || ctx->ws_bytes < need)
    return gcnb_fail(ctx, GCNB_E_WORKSPACE, "tcgen05 highway needs %s%lld workspace bytes, have %lld", "",
                     (long long)need, (long long)ctx->ws_bytes);
  const int ldw = ld32(hd);
  const size_t wsz = (size_t)hd * ldw;
  float* WhT = reinterpret_cast<float*>(ctx->ws);
  float* WhL = WhT + wsz;
  float* WtT = WhL + wsz;
  float* WtL = WtT + wsz;
  int rc = split_weights(ctx, Wh, ldwh, hd, hd, 1, WhT, WhL, ldw);
  if (rc != GCNB_OK) return rc;
  rc = split_weights(ctx, Wt, ldwt, hd, hd, 1, WtT, WtL, ldw);
  if (rc != GCNB_OK) return rc;
  TcParams p;
  memset(&p, 0, sizeof(p));
  const Tiling tl = pick_tiling(ctx, hd, 1, &p);
  if (!make_opmap(tl, &p.mapA[0], S, n_rows, hd, lds, BM) || !make_opmap(tl, &p.mapB[0], WhT, hd, hd, ldw, p.BN) ||
      !make_opmap(tl, &p.mapBlo[0], WhL, hd, hd, ldw, p.BN) || !make_opmap(tl, &p.mapA[1], X, n_rows, hd, ldx, BM) ||
      !make_opmap(tl, &p.mapB[1], WtT, hd, hd, ldw, p.BN) || !make_opmap(tl, &p.mapBlo[1], WtL, hd, hd, ldw, p.BN))
    return gcnb_fail(ctx, GCNB_E_CUDA, "cuTensorMapEncodeTiled failed%s", "");
  if (!make_map(&p.mapIn, X, n_rows, hd, ldx, BM) || !make_map(&p.mapOut[0], Y, n_rows, hd, ldy, BM) ||
      (H && !make_map(&p.mapOut[1], H, n_rows, hd, ldh, BM)) || (T && !make_map(&p.mapOut[2], T, n_rows, hd, ldt, BM)))
    return gcnb_fail(ctx, GCNB_E_CUDA, "cuTensorMapEncodeTiled failed%s", "");
  p.nphase = 2;
  p.hw = 1;
  p.kblocks[0] = p.kblocks[1] = cdiv(hd, tl.bk);
  p.M = n_rows; p.N = hd;
  p.C = Y; p.ldc = ldy; p.bias = bh; p.act = act; p.accumulate = 0;
  p.bias_t = bt; p.X = X; p.ldx = ldx; p.H = H; p.ldh = ldh; p.T = T; p.ldt = ldt;
  p.dbg = reinterpret_cast<long long*>(ctx->tc_dbg);
  p.dbg_mode = ctx->tc_dbg_mode;
  p.blo_in_kernel = ctx->gemm_v >= 2 ? ctx->gemm_blo2 : ctx->gemm_blo;
  return launch(ctx, p);
}

// ---------------------------------------------------------------------------------------------
namespace {
struct WgPlan { int BN, m_tiles, n_tiles, splits, kb_per_split; };
// slots = CTAs the grid may hold at once (SMs x CTAs per SM); bk = k-rows per stage
WgPlan wgrad_plan(int slots, int M, int N, int K, int bk) {
  WgPlan w;
  pick_bn(N, &w.BN, &w.n_tiles);
  w.m_tiles = (M + BM - 1) / BM;
  const int tiles = w.m_tiles * w.n_tiles;
  const int kb_total = (K + bk - 1) / bk;
  int splits = slots / tiles;
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
  if (M > 4096 || N > 4096) return false;  // weight-shaped outputs only (the hot-column block has up to 4096 rows)
  if ((lda % 4) != 0 || (ldb % 4) != 0) return false;
  return encode_fn() != nullptr;
}

size_t gcnb_wgrad_tc_workspace_bytes(int M, int N, int K) {
  // no context here (the ABI sizes workspaces before a context exists): bound the split count by twice the largest SM
  // count any sm_100 part has (two CTAs per SM) instead of assuming 148; the launch plans with ctx->sm_count
  const int slots = 512;
  const WgPlan w = wgrad_plan(slots, M, N, K, BK2);
  const int tiles = w.m_tiles * w.n_tiles;
  int max_splits = tiles > 0 ? slots / tiles : 1;
  if (max_splits < w.splits) max_splits = w.splits;
  if (max_splits < 1) max_splits = 1;
  return (size_t)(max_splits + 1) * M * N * sizeof(float);
}

int gcnb_wgrad_tc(gcnb_ctx* ctx, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                  int ldc, int accumulate) {
  GCNB_REQUIRE(ctx, aligned16(A) && aligned16(B), "tcgen05 wgrad: 16-byte aligned matrices");
  const bool v2 = ctx->gemm_v >= 2;
  const int bk = v2 ? BK2 : BK;
  const WgPlan w = wgrad_plan(ctx->sm_count * (v2 ? 2 : 1), M, N, K, bk);
  const size_t need = (size_t)w.splits * M * N * sizeof(float);
  if (!ctx->ws || ctx->ws_bytes < need)
    return gcnb_fail(ctx, GCNB_E_WORKSPACE, "tcgen05 wgrad needs %s%lld workspace bytes, have %lld", "",
                     (long long)need, (long long)ctx->ws_bytes);
  WgParams p;
  memset(&p, 0, sizeof(p));
  if (!make_map(&p.mapA, A, K, M, lda, bk, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
      !make_map(&p.mapB, B, K, N, ldb, bk, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
    return gcnb_fail(ctx, GCNB_E_CUDA, "cuTensorMapEncodeTiled failed%s", "");
  p.M = M; p.N = N; p.K = K; p.BN = w.BN; p.m_tiles = w.m_tiles; p.n_tiles = w.n_tiles;
  p.splits = w.splits; p.kb_per_split = w.kb_per_split;
  p.partial = reinterpret_cast<float*>(ctx->ws);
  // a stage holds (BM + BN) / 32 chunks of bk x 128 B, raw and residual
  const size_t smem = (size_t)kStages * 2 * (BM / 32 + p.BN / 32) * bk * 128 + (3 * kStages + 2) * 8 + 32 + 1024;
  const unsigned grid = (unsigned)(w.m_tiles * w.n_tiles * w.splits);
  if (v2) {
    GCNB_CUDA(ctx, cudaFuncSetAttribute(wgrad_tc_kernel<BK2, kTmemCols2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_tc_kernel<BK2, kTmemCols2, 2><<<grid, kThreads, smem, ctx->stream>>>(p);
  } else {
    GCNB_CUDA(ctx, cudaFuncSetAttribute(wgrad_tc_kernel<BK, kTmemCols, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_tc_kernel<BK, kTmemCols, 1><<<grid, kThreads, smem, ctx->stream>>>(p);
  }
  GCNB_LAUNCHED(ctx);
  ctx->tc_launches++;
  const long long n = (long long)M * N;
  wgrad_reduce_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(p.partial, w.splits, M, N, C, ldc, accumulate);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}
