// spmm.cu -- CSR x dense SpMM for sm_100a:  C = epilogue(A . B).
//
// Replaces theano.sparse.structured_dot (reference gcnmodel.py:39,130,153 and its gradient): the
// hot loop of the GCN (A_hat.H), the first-layer projection (X.W0) and the transposed product
// X^T.dz all run through the kernels in this file.
//
// Work decomposition: the host plan (gcnb_csr_plan) cuts every row into "items" of at most
// `chunk` nonzeros.  One warp owns one item: lanes are spread across the feature columns
// (float4 per lane, NCHUNK float4s per lane => up to 512 columns per pass), the warp walks the
// item's nonzeros in CSR order and accumulates val * B[col, :] in registers -- a segmented,
// order-preserving reduction over the nonzeros of the row.  Single-item rows run the fused
// epilogue and store once; multi-item ("long") rows store per-item partial sums that
// spmm_fixup_kernel adds in item order, so results do not depend on scheduling.
//
// Two gather engines, selected per context ("spmm_variant"):
//   0  LDG.128 register gather: U nonzeros x NCHUNK 128-bit loads in flight per warp.
//   1  bulk-copy (TMA engine, cp.async.bulk -> SASS UBLKCP) staged gather: each warp owns a ring
//      of shared-memory row slots armed with mbarriers; one lane issues a 1-D bulk copy per
//      gathered row, the warp consumes rows from shared memory with conflict-free LDS.128.
//      Persistent CTAs (one per SM) pull items from an atomic counter.
//   2  L2-resident column panels: the product is computed one PW-column panel (PW = 16/32/64 floats) of B and
//      C at a time, panel-major over the whole grid (blockIdx.y = panel).  A panel of B is rows x PW*4 bytes
//      (N = 500k, PW = 32: 64 MB), so after its first touch every further gather of the panel is an L2 hit:
//      B leaves HBM once per product instead of once per nonzero.  A warp is cut into 32/(PW/4) lane groups,
//      each group owns one row item and walks its nonzeros in CSR order (same order-preserving sum).
//
// HBM model: per nonzero one K-wide fp32 row of B is read (K*4 B) plus 8 B of CSR;
// per row K*4 B are written.
#include "common.cuh"

namespace {

struct SpmmParams {
  const int4* items;
  int n_items;
  const int* col;
  const float* val;
  const float* B;
  int ldb;
  float* C;
  int ldc;
  int K;
  int col0;  // first column of this pass (multiple of 4)
  int nf4;   // float4s per row in this pass
  float* partial;
  int ldp;  // floats per partial slot
  const float* bias;
  int act;
  int softmax;
  int accumulate;
  uint32_t thresh;  // dropout keep threshold (0 => no dropout)
  float scale;
  uint64_t seed;
  int64_t row0;
  float* logits;
  const int* long_rows;
  int n_long;
  int* counter;
  int pcol0;       // first column of this pass inside a partial slot (0 unless the slots span all of K: panel engine)
  int k4;          // K rounded up to 4 (panel engine: the panels tile [0, k4))
  int evict_last;  // panel engine, gather cache policy: 0 default, 1 L2 evict_last hint, 2 the same with L1 allocation
  int col_base;    // panel engine: first column of panel 0 of this launch
  // feature-sliced product of a row-partitioned run (gcnb_spmm_csr_sliced_f32): B holds this rank's column slice, the
  // columns computed here are out_col0 + [0, k4) of the full operand, and row i of the result belongs to rank i / n_pad
  int out_col0;
  int n_pad;       // 0: every row is stored through C
  float* peer_C[GCNB_MAX_PEERS];
  // fused push (gcnb_push_arm): the finished output row block also goes to the owners of its columns, as the operand
  // of the feature-sliced graph convolution that follows (first layer: H0 = dropout(act(X.W0 + b0)))
  PushPlan push;
};

// first element of output row `row`: local C, or the owning rank's copy of C in a feature-sliced product
__device__ __forceinline__ float* out_row(const SpmmParams& p, int row) {
  if (p.n_pad > 0) {
    const int q = row / p.n_pad;
    return p.peer_C[q] + (size_t)(row - q * p.n_pad) * p.ldc;
  }
  return p.C + (size_t)row * p.ldc;
}

constexpr int kWarpsPerCta = 8;

// ---------------------------------------------------------------------------------------
// fused epilogue: the warp holds one output row (acc[ch] = float4 number lane + 32*ch)
// ---------------------------------------------------------------------------------------
template <int NCHUNK>
__device__ __forceinline__ void spmm_epilogue(const SpmmParams& p, int row, float4 (&acc)[NCHUNK], int lane) {
  float v[NCHUNK][4];
#pragma unroll
  for (int ch = 0; ch < NCHUNK; ++ch) {
    v[ch][0] = acc[ch].x; v[ch][1] = acc[ch].y; v[ch][2] = acc[ch].z; v[ch][3] = acc[ch].w;
  }
  const int gcol0 = p.col0 + p.out_col0;  // first column of this pass in the full operand
  float* const orow = out_row(p, row);
  if (p.accumulate == 2) {  // pre-activation accumulate: the product joins what C already holds (dense hot-column part)
    const float4* crow_in = reinterpret_cast<const float4*>(orow + gcol0);
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
      const int f4 = lane + 32 * ch;
      if (f4 < p.nf4) {
        const float4 c = crow_in[f4];
        v[ch][0] += c.x; v[ch][1] += c.y; v[ch][2] += c.z; v[ch][3] += c.w;
      }
    }
  }
  if (p.bias != nullptr) {
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
      const int f4 = lane + 32 * ch;
      if (f4 < p.nf4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + gcol0) + f4);
        v[ch][0] += b.x; v[ch][1] += b.y; v[ch][2] += b.z; v[ch][3] += b.w;
      }
    }
  }
  if (p.softmax) {
    if (p.logits != nullptr) {
#pragma unroll
      for (int ch = 0; ch < NCHUNK; ++ch) {
        const int f4 = lane + 32 * ch;
        if (f4 < p.nf4) {
          float4 o;
          const int c = gcol0 + 4 * f4;
          o.x = c + 0 < p.K ? v[ch][0] : 0.f; o.y = c + 1 < p.K ? v[ch][1] : 0.f;
          o.z = c + 2 < p.K ? v[ch][2] : 0.f; o.w = c + 3 < p.K ? v[ch][3] : 0.f;
          reinterpret_cast<float4*>(p.logits + (size_t)row * p.ldc + gcol0)[f4] = o;
        }
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
      const int c = gcol0 + 4 * (lane + 32 * ch);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (lane + 32 * ch < p.nf4 && c + e < p.K) m = fmaxf(m, v[ch][e]);
    }
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
      const int c = gcol0 + 4 * (lane + 32 * ch);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = lane + 32 * ch < p.nf4 && c + e < p.K;
        v[ch][e] = ok ? expf(v[ch][e] - m) : 0.f;
        s += v[ch][e];
      }
    }
    s = warp_sum(s);
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch)
#pragma unroll
      for (int e = 0; e < 4; ++e) v[ch][e] = v[ch][e] / s;
  } else {
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
      const int f4 = lane + 32 * ch;
      const int c = gcol0 + 4 * f4;
      if (f4 < p.nf4) {
        if (p.act != GCNB_ACT_LINEAR) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[ch][e] = act_apply(p.act, v[ch][e]);
        }
        if (p.thresh != 0u) {
          const uint4 r = dropout_draw(p.seed, p.row0 + row, (uint32_t)(c >> 2));
          v[ch][0] = r.x < p.thresh ? v[ch][0] * p.scale : 0.f;
          v[ch][1] = r.y < p.thresh ? v[ch][1] * p.scale : 0.f;
          v[ch][2] = r.z < p.thresh ? v[ch][2] * p.scale : 0.f;
          v[ch][3] = r.w < p.thresh ? v[ch][3] * p.scale : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c + e >= p.K) v[ch][e] = 0.f;  // padding columns stay zero
      }
    }
  }
  float4* crow = reinterpret_cast<float4*>(orow + gcol0);
#pragma unroll
  for (int ch = 0; ch < NCHUNK; ++ch) {
    const int f4 = lane + 32 * ch;
    if (f4 < p.nf4) {
      float4 o = make_float4(v[ch][0], v[ch][1], v[ch][2], v[ch][3]);
      if (p.accumulate == 1) {
        const float4 old = crow[f4];
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      st_stream_f4(crow + f4, o);
      if (p.push.on) push_store_f4(p.push, row, gcol0 + 4 * f4, o);
    }
  }
}

template <int NCHUNK>
__device__ __forceinline__ void spmm_store_item(const SpmmParams& p, const int4 it, float4 (&acc)[NCHUNK],
                                                int lane) {
  if (it.w < 0) {
    spmm_epilogue<NCHUNK>(p, it.x, acc, lane);
  } else {
    float4* dst = reinterpret_cast<float4*>(p.partial + (size_t)it.w * p.ldp + p.pcol0);
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch)
      if (lane + 32 * ch < p.nf4) dst[lane + 32 * ch] = acc[ch];
  }
}

// ---------------------------------------------------------------------------------------
// variant 0: LDG.128 register gather
// ---------------------------------------------------------------------------------------
template <int NCHUNK, int U>
__global__ void __launch_bounds__(kWarpsPerCta * 32) spmm_ldg_kernel(const SpmmParams p) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (item >= p.n_items) return;
  const int4 it = __ldg(p.items + item);
  const uint64_t pol = policy_evict_first();
  float4 acc[NCHUNK];
#pragma unroll
  for (int ch = 0; ch < NCHUNK; ++ch) acc[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* Bs = reinterpret_cast<const float4*>(p.B + p.col0) + lane;
  const size_t ldb4 = (size_t)(p.ldb >> 2);

  for (int base = it.y; base < it.z; base += 32) {
    const int k = base + lane;
    int c = 0;
    float a = 0.f;
    if (k < it.z) {
      c = ld_stream_s32(p.col + k, pol);
      a = ld_stream_f32(p.val + k, pol);
    }
    const int cnt = min(32, it.z - base);
    for (int j = 0; j < cnt; j += U) {
      float4 x[U][NCHUNK];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cu = __shfl_sync(0xffffffffu, c, (j + u) & 31);
        const float4* src = Bs + (size_t)cu * ldb4;
        const bool live = (j + u) < cnt;
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch) {
          if (live && lane + 32 * ch < p.nf4) x[u][ch] = ld_gather_f4(src + 32 * ch);
          else x[u][ch] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float au = __shfl_sync(0xffffffffu, a, (j + u) & 31);
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch) {
          acc[ch].x = fmaf(au, x[u][ch].x, acc[ch].x);
          acc[ch].y = fmaf(au, x[u][ch].y, acc[ch].y);
          acc[ch].z = fmaf(au, x[u][ch].z, acc[ch].z);
          acc[ch].w = fmaf(au, x[u][ch].w, acc[ch].w);
        }
      }
    }
  }
  spmm_store_item<NCHUNK>(p, it, acc, lane);
}

// ---------------------------------------------------------------------------------------
// variant 1: bulk-copy (TMA engine) staged gather, persistent CTAs
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

constexpr int kBulkGroup = 4;  // gathered rows per mbarrier phase

template <int NCHUNK, int STAGES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) spmm_bulk_kernel(const SpmmParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)p.nf4 * 16u;
  const uint32_t slot_bytes = (row_bytes + 127u) & ~127u;
  const uint32_t stage_bytes = slot_bytes * kBulkGroup;
  unsigned char* my_ring = smem + (size_t)warp * STAGES * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * STAGES * stage_bytes) + warp * STAGES;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint32_t ring_u32 = smem_u32(my_ring);
  const uint32_t bars_u32 = smem_u32(bars);
  const char* Bbytes = reinterpret_cast<const char*>(p.B + p.col0);
  const size_t ldb_bytes = (size_t)p.ldb * 4;

  const uint64_t pol = policy_evict_first();
  uint32_t gseq = 0;  // groups issued == groups consumed at item boundaries (ring position)
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(p.counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= p.n_items) break;
    const int4 it = __ldg(p.items + item);
    const int n = it.z - it.y;
    const int ngroups = (n + kBulkGroup - 1) / kBulkGroup;
    float4 acc[NCHUNK];
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) acc[ch] = make_float4(0.f, 0.f, 0.f, 0.f);

    int ci = 0;      // column block of the issue side (32 nonzeros)
    float vc = 0.f;  // value block of the consume side
    int issued = 0;
    // issue group g (all lanes execute; lanes < kBulkGroup issue one row copy each)
    auto issue = [&](int g) {
      const int k0 = g * kBulkGroup;  // nonzero offset inside the item
      if ((k0 & 31) == 0) {
        const int k = it.y + k0 + lane;
        ci = k < it.z ? ld_stream_s32(p.col + k, pol) : 0;
      }
      const int rows = min(kBulkGroup, n - k0);
      const uint32_t stage = (gseq + (uint32_t)g) % STAGES;
      const uint32_t bar = bars_u32 + stage * 8u;
      const int c = __shfl_sync(0xffffffffu, ci, (k0 + lane) & 31);
      if (lane == 0) mbar_expect_tx(bar, row_bytes * (uint32_t)rows);
      __syncwarp();
      if (lane < rows)
        bulk_g2s(ring_u32 + stage * stage_bytes + (uint32_t)lane * slot_bytes, Bbytes + (size_t)c * ldb_bytes,
                 row_bytes, bar);
    };
    for (; issued < ngroups && issued < STAGES; ++issued) issue(issued);

    for (int g = 0; g < ngroups; ++g) {
      const int k0 = g * kBulkGroup;
      if ((k0 & 31) == 0) {
        const int k = it.y + k0 + lane;
        vc = k < it.z ? ld_stream_f32(p.val + k, pol) : 0.f;
      }
      const uint32_t seq = gseq + (uint32_t)g;
      const uint32_t stage = seq % STAGES;
      mbar_wait(bars_u32 + stage * 8u, (seq / STAGES) & 1u);
      const int rows = min(kBulkGroup, n - k0);
      const unsigned char* sbase = my_ring + (size_t)stage * stage_bytes;
#pragma unroll
      for (int r = 0; r < kBulkGroup; ++r) {
        if (r < rows) {
          const float a = __shfl_sync(0xffffffffu, vc, (k0 + r) & 31);
          const float4* srow = reinterpret_cast<const float4*>(sbase + (size_t)r * slot_bytes);
#pragma unroll
          for (int ch = 0; ch < NCHUNK; ++ch) {
            if (lane + 32 * ch < p.nf4) {
              const float4 x = srow[lane + 32 * ch];
              acc[ch].x = fmaf(a, x.x, acc[ch].x);
              acc[ch].y = fmaf(a, x.y, acc[ch].y);
              acc[ch].z = fmaf(a, x.z, acc[ch].z);
              acc[ch].w = fmaf(a, x.w, acc[ch].w);
            }
          }
        }
      }
      __syncwarp();  // every lane has read the stage before it is re-armed
      if (issued < ngroups) {
        issue(issued);
        ++issued;
      }
    }
    gseq += (uint32_t)ngroups;
    spmm_store_item<NCHUNK>(p, it, acc, lane);
  }
}

// ---------------------------------------------------------------------------------------
// variant 2: L2-resident column panels
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// same with L1 allocation (gathered operands with reuse inside an SM: the frequent terms of X.W0)
__device__ __forceinline__ float4 ld_gather_hint_l1_f4(const float4* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float4 ld_gather_hint_f4(const float4* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}

// element-wise epilogue of one float4 of an output row (no row softmax: that needs the whole row, see
// row_softmax_kernel).  Same arithmetic, in the same order, as spmm_epilogue.
__device__ __forceinline__ void panel_epilogue(const SpmmParams& p, int row, int c_local, float4 acc) {
  float v[4] = {acc.x, acc.y, acc.z, acc.w};
  const int c = c_local + p.out_col0;  // column in the full operand (bias, padding rule, dropout counter, store)
  float4* cptr = reinterpret_cast<float4*>(out_row(p, row) + c);
  if (p.accumulate == 2) {
    const float4 o = *cptr;
    v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
  }
  if (p.bias != nullptr) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c));
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (p.act != GCNB_ACT_LINEAR) {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = act_apply(p.act, v[e]);
  }
  if (p.thresh != 0u) {
    const uint4 r = dropout_draw(p.seed, p.row0 + row, (uint32_t)(c >> 2));
    v[0] = r.x < p.thresh ? v[0] * p.scale : 0.f;
    v[1] = r.y < p.thresh ? v[1] * p.scale : 0.f;
    v[2] = r.z < p.thresh ? v[2] * p.scale : 0.f;
    v[3] = r.w < p.thresh ? v[3] * p.scale : 0.f;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (c + e >= p.K) v[e] = 0.f;  // padding columns stay zero
  float4 o = make_float4(v[0], v[1], v[2], v[3]);
  if (p.accumulate == 1) {
    const float4 old = *cptr;
    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
  }
  st_stream_f4(cptr, o);
  if (p.push.on) push_store_f4(p.push, row, c, o);
}

// GL lanes per row item (panel = 4*GL floats), R (column, value) pairs per lane and batch: GL*R gathers in flight.
// KEEP: gather cache policy (0 default, 1 L2 evict_last hint, 2 the same with L1 allocation).  Batches in which every item of the warp still has GL*R nonzeros run
// without predicates (plans sorted by item length make that the common case); the rest take the predicated tail.
template <int GL, int R, int KEEP>
__global__ void __launch_bounds__(kWarpsPerCta * 32, GL * R <= 8 ? 4 : 2) spmm_panel_kernel(const SpmmParams p) {
  constexpr int G = 32 / GL;
  constexpr int NB = GL * R;
  const int lane = threadIdx.x & 31;
  const int g = lane / GL, gl = lane % GL;
  const int col0 = p.col_base + (int)blockIdx.y * (GL * 4);
  const int item = (blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5)) * G + g;
  const bool valid = item < p.n_items;
  int4 it = make_int4(0, 0, 0, -1);
  if (valid) it = __ldg(p.items + item);
  const int n = it.z - it.y;
  const int nmax = __reduce_max_sync(0xffffffffu, n);
  const int nmin = __reduce_min_sync(0xffffffffu, n);
  const bool lane_on = col0 + 4 * gl < p.k4;
  const uint64_t pol = policy_evict_first();
  const uint64_t keep = policy_evict_last();
  // lanes past the last column of a partial panel gather the panel's first float4 instead (in bounds) and never store
  const char* Bs = reinterpret_cast<const char*>(p.B + col0 + (lane_on ? 4 * gl : 0));
  const uint32_t row_bytes = (uint32_t)p.ldb * 4u;
  const int* colp = p.col + it.y + gl;
  const float* valp = p.val + it.y + gl;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

  auto gather = [&](int cj) -> float4 {
    const float4* src = reinterpret_cast<const float4*>(Bs + (uint64_t)(uint32_t)cj * row_bytes);
    return KEEP == 2 ? ld_gather_hint_l1_f4(src, keep) : KEEP == 1 ? ld_gather_hint_f4(src, keep) : ld_gather_f4(src);
  };

  int base = 0;
  if (NB <= nmin) {  // batches every lane group can fill: no predicates; (column, value) pairs fetched one batch ahead
    int c[R], cn[R];
    float a[R], an[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      c[r] = ld_stream_s32(colp + r * GL, pol);
      a[r] = ld_stream_f32(valp + r * GL, pol);
      cn[r] = 0;
      an[r] = 0.f;
    }
    for (; base + NB <= nmin; base += NB) {
      if (base + 2 * NB <= nmin) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          cn[r] = ld_stream_s32(colp + base + NB + r * GL, pol);
          an[r] = ld_stream_f32(valp + base + NB + r * GL, pol);
        }
      }
      float4 x[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) x[j] = gather(__shfl_sync(0xffffffffu, c[j / GL], j % GL, GL));
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float aj = __shfl_sync(0xffffffffu, a[j / GL], j % GL, GL);
        acc.x = fmaf(aj, x[j].x, acc.x);
        acc.y = fmaf(aj, x[j].y, acc.y);
        acc.z = fmaf(aj, x[j].z, acc.z);
        acc.w = fmaf(aj, x[j].w, acc.w);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        c[r] = cn[r];
        a[r] = an[r];
      }
    }
  }
  for (; base < nmax; base += NB) {  // ragged tail: per-item predicates
    int c[R];
    float a[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int k = base + r * GL + gl;
      c[r] = 0;
      a[r] = 0.f;
      if (k < n) {
        c[r] = ld_stream_s32(colp + base + r * GL, pol);
        a[r] = ld_stream_f32(valp + base + r * GL, pol);
      }
    }
    float4 x[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int cj = __shfl_sync(0xffffffffu, c[j / GL], j % GL, GL);
      x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (base + j < n) x[j] = gather(cj);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const float aj = __shfl_sync(0xffffffffu, a[j / GL], j % GL, GL);
      acc.x = fmaf(aj, x[j].x, acc.x);
      acc.y = fmaf(aj, x[j].y, acc.y);
      acc.z = fmaf(aj, x[j].z, acc.z);
      acc.w = fmaf(aj, x[j].w, acc.w);
    }
  }
  if (!valid || !lane_on) return;
  if (it.w >= 0)
    *reinterpret_cast<float4*>(p.partial + (size_t)it.w * p.ldp + col0 + 4 * gl) = acc;
  else
    panel_epilogue(p, it.x, col0 + 4 * gl, acc);
}

// row softmax in place over C (panel engine: the product and bias are already in C); optional copy of the logits
template <int NCHUNK>
__global__ void __launch_bounds__(kWarpsPerCta * 32) row_softmax_kernel(const SpmmParams p, int n_rows) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  float4 acc[NCHUNK];
  const float4* crow = reinterpret_cast<const float4*>(p.C + (size_t)row * p.ldc);
#pragma unroll
  for (int ch = 0; ch < NCHUNK; ++ch)
    acc[ch] = lane + 32 * ch < p.nf4 ? crow[lane + 32 * ch] : make_float4(0.f, 0.f, 0.f, 0.f);
  spmm_epilogue<NCHUNK>(p, row, acc, lane);
}

// row softmax in place over rows wider than one register pass (K > 512: TwitterWorld has 930 classes at bucket 2400,
// reference README.md:177-181).  C already holds product + bias; three sweeps over the row (max, sum of exp,
// normalise), the row stays in L1/L2 between them.  Padding columns are written as zeros.
__global__ void __launch_bounds__(kWarpsPerCta * 32) row_softmax_wide_kernel(const SpmmParams p, int n_rows) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  float4* crow = reinterpret_cast<float4*>(p.C + (size_t)row * p.ldc);
  float4* lrow = p.logits ? reinterpret_cast<float4*>(p.logits + (size_t)row * p.ldc) : nullptr;
  const int nf4 = p.k4 >> 2;
  float m = -INFINITY;
  for (int f4 = lane; f4 < nf4; f4 += 32) {
    float4 x = crow[f4];
    const int c = 4 * f4;
    if (c + 1 >= p.K) x.y = 0.f;
    if (c + 2 >= p.K) x.z = 0.f;
    if (c + 3 >= p.K) x.w = 0.f;
    if (lrow) lrow[f4] = x;
    m = fmaxf(m, x.x);
    if (c + 1 < p.K) m = fmaxf(m, x.y);
    if (c + 2 < p.K) m = fmaxf(m, x.z);
    if (c + 3 < p.K) m = fmaxf(m, x.w);
  }
  m = warp_max(m);
  float s = 0.f;
  for (int f4 = lane; f4 < nf4; f4 += 32) {
    const float4 x = crow[f4];
    const int c = 4 * f4;
    s += expf(x.x - m);
    if (c + 1 < p.K) s += expf(x.y - m);
    if (c + 2 < p.K) s += expf(x.z - m);
    if (c + 3 < p.K) s += expf(x.w - m);
  }
  s = warp_sum(s);
  for (int f4 = lane; f4 < nf4; f4 += 32) {
    const float4 x = crow[f4];
    const int c = 4 * f4;
    float4 o;
    o.x = expf(x.x - m) / s;
    o.y = c + 1 < p.K ? expf(x.y - m) / s : 0.f;
    o.z = c + 2 < p.K ? expf(x.z - m) / s : 0.f;
    o.w = c + 3 < p.K ? expf(x.w - m) / s : 0.f;
    crow[f4] = o;
  }
}

// ---------------------------------------------------------------------------------------
// long rows: add the per-item partial sums in item order, then the epilogue
// ---------------------------------------------------------------------------------------
template <int NCHUNK>
__global__ void __launch_bounds__(kWarpsPerCta * 32) spmm_fixup_kernel(const SpmmParams p) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (i >= p.n_long) return;
  const int row = p.long_rows[3 * i], s0 = p.long_rows[3 * i + 1], ns = p.long_rows[3 * i + 2];
  float4 acc[NCHUNK];
#pragma unroll
  for (int ch = 0; ch < NCHUNK; ++ch) acc[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = s0; s < s0 + ns; ++s) {
    const float4* src = reinterpret_cast<const float4*>(p.partial + (size_t)s * p.ldp + p.pcol0);
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
      if (lane + 32 * ch < p.nf4) {
        const float4 x = src[lane + 32 * ch];
        acc[ch].x += x.x; acc[ch].y += x.y; acc[ch].z += x.z; acc[ch].w += x.w;
      }
    }
  }
  spmm_epilogue<NCHUNK>(p, row, acc, lane);
}

template <int NCHUNK>
int launch_pass(gcnb_ctx* ctx, const SpmmParams& p, int engine, int unroll) {
  const int grid = cdiv(p.n_items, kWarpsPerCta);
  if (p.n_items > 0) {
    if (engine == 1) {
      constexpr int STAGES = NCHUNK == 4 ? 3 : 4, WARPS = 8;
      const uint32_t row_bytes = (uint32_t)p.nf4 * 16u;
      const uint32_t slot_bytes = (row_bytes + 127u) & ~127u;
      const size_t smem = (size_t)WARPS * STAGES * kBulkGroup * slot_bytes + (size_t)WARPS * STAGES * 8;
      auto kern = spmm_bulk_kernel<NCHUNK, STAGES, WARPS>;
      GCNB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      GCNB_CUDA(ctx, cudaMemsetAsync(p.counter, 0, sizeof(int), ctx->stream));
      int ctas_per_sm = (int)((220u * 1024u) / (smem + 1024));
      if (ctas_per_sm < 1) ctas_per_sm = 1;
      if (ctas_per_sm > 4) ctas_per_sm = 4;
      int sms = ctx->sm_count - ctx->sm_margin;
      if (sms < 1) sms = 1;
      int g = sms * ctas_per_sm;
      if (g > grid) g = grid;
      kern<<<g, WARPS * 32, smem, ctx->stream>>>(p);
      GCNB_LAUNCHED(ctx);
    } else {
      int U = unroll;
      if (U == 0) U = NCHUNK <= 2 ? 8 : 2;
      if (U >= 8) spmm_ldg_kernel<NCHUNK, 8><<<grid, kWarpsPerCta * 32, 0, ctx->stream>>>(p);
      else if (U >= 4) spmm_ldg_kernel<NCHUNK, 4><<<grid, kWarpsPerCta * 32, 0, ctx->stream>>>(p);
      else spmm_ldg_kernel<NCHUNK, 2><<<grid, kWarpsPerCta * 32, 0, ctx->stream>>>(p);
      GCNB_LAUNCHED(ctx);
    }
  }
  if (p.n_long > 0) {
    spmm_fixup_kernel<NCHUNK><<<cdiv(p.n_long, kWarpsPerCta), kWarpsPerCta * 32, 0, ctx->stream>>>(p);
    GCNB_LAUNCHED(ctx);
  }
  return GCNB_OK;
}

constexpr int kMaxPassCols = 512;

// Gather engine of one product.  A->engine >= 0 names it, -1 defers to the context option, -2 chooses from the
// shapes (measured on B200, C3 shapes, profiles/r1c_spmm_panel_sweep.txt and r1_spmm_sweep.txt):
//  * the dense operand is larger than L2 can trivially hold (> 32 MB) but one 32-column panel of it
//    (n_cols x 128 B) fits (<= 80 MB): L2-resident column panels -- A_hat.H at K=300 2.30 ms vs 3.37 ms for the
//    best HBM gather, K=256 1.70 vs 2.34 ms, X_cold.W0 5.74 vs 7.51 ms;
//  * else, operand > 96 MB and rows wider than 64 float4 lanes: bulk-copy staged gather (91% of the HBM copy rate
//    vs 71-78% for register gathers);
//  * else (L2-resident operands, narrow rows): LDG register gather, 2 nonzeros per batch.
void pick_engine(const gcnb_ctx* ctx, const gcnb_csr* A, int ldb, int K4, int* engine, int* unroll) {
  *engine = A->engine >= 0 ? A->engine : ctx->spmm_variant;
  *unroll = A->unroll > 0 ? A->unroll : ctx->spmm_unroll;
  if (A->engine == -2) {
    const size_t operand_bytes = (size_t)A->n_cols * (size_t)ldb * sizeof(float);
    const size_t panel_bytes = (size_t)A->n_cols * 128;
    // (rows that average fewer than 4 nonzeros -- the alpha = 1.5 power-law graph is 500k self loops plus a few hubs --
    // are bound by the per-item latency chain, which the panel engine pays once per panel: 0.74 ms for 538k nonzeros)
    const bool very_sparse = A->nnz < 4 * (int64_t)A->n_rows;
    if (operand_bytes > ((size_t)32 << 20) && panel_bytes <= ((size_t)80 << 20) && K4 >= 64 && !very_sparse) *engine = 2;
    else *engine = (operand_bytes > ((size_t)96 << 20) && K4 > 256 && !very_sparse) ? 1 : 0;
    if (*unroll == 0 && *engine != 2) *unroll = 2;
  }
}

template <int NCHUNK>
void launch_fixup(gcnb_ctx* ctx, const SpmmParams& p) {
  spmm_fixup_kernel<NCHUNK><<<cdiv(p.n_long, kWarpsPerCta), kWarpsPerCta * 32, 0, ctx->stream>>>(p);
}
template <int NCHUNK>
void launch_softmax(gcnb_ctx* ctx, const SpmmParams& p, int n_rows) {
  row_softmax_kernel<NCHUNK><<<cdiv(n_rows, kWarpsPerCta), kWarpsPerCta * 32, 0, ctx->stream>>>(p, n_rows);
}

// engine 2: every column panel of the product in one launch (panel-major), then the long-row fix-up and, when
// asked for, the row softmax as a pass of its own
int launch_panels(gcnb_ctx* ctx, const SpmmParams& p0, int n_rows, int K4, int unroll, bool auto_width) {
  int PW = ctx->spmm_panel;
  if (PW != 16 && PW != 64) PW = 32;
  if (auto_width) PW = 32;
  SpmmParams p = p0;
  p.softmax = 0;
  p.logits = nullptr;
  p.col0 = 0;
  p.k4 = K4;
  p.nf4 = K4 / 4;
  p.ldp = K4;
  p.evict_last = ctx->spmm_panel_policy;
  if (p.n_items > 0) {
    const int threads = kWarpsPerCta * 32;
#define GCNB_PANEL(GLV, RV, NPANELS)                                                                      \
  do {                                                                                                    \
    const dim3 grid(cdiv(p.n_items, kWarpsPerCta * (32 / GLV)), NPANELS);                                 \
    if (p.evict_last == 2) spmm_panel_kernel<GLV, RV, 2><<<grid, threads, 0, ctx->stream>>>(p);           \
    else if (p.evict_last) spmm_panel_kernel<GLV, RV, 1><<<grid, threads, 0, ctx->stream>>>(p);           \
    else spmm_panel_kernel<GLV, RV, 0><<<grid, threads, 0, ctx->stream>>>(p);                             \
    GCNB_LAUNCHED(ctx);                                                                                   \
  } while (0)
    if (PW == 32) {
      // full 32-column panels, then a remainder of at most 16 columns on 4-lane groups (8 items per warp): K = 300
      // is 9 panels + 12 columns, and a tenth 8-lane pass would spend a full pass on 12 useful columns
      const int full = K4 / 32, rem = K4 - 32 * full;
      const int n32 = rem > 16 ? full + 1 : full;
      if (n32 > 0) {
        if (unroll >= 16) GCNB_PANEL(8, 2, n32);
        else GCNB_PANEL(8, 1, n32);
      }
      if (rem > 0 && rem <= 16) {
        p.col_base = 32 * full;
        GCNB_PANEL(4, 2, 1);
        p.col_base = 0;
      }
    } else if (PW == 16) {
      if (unroll >= 16) GCNB_PANEL(4, 4, cdiv(K4, 16));
      else GCNB_PANEL(4, 2, cdiv(K4, 16));
    } else {
      GCNB_PANEL(16, 1, cdiv(K4, 64));
    }
#undef GCNB_PANEL
  }
  for (int pass = 0; pass < 2; ++pass) {  // 0: long-row fix-up, 1: row softmax
    if (pass == 0 && p.n_long == 0) continue;
    if (pass == 1 && !p0.softmax) continue;
    SpmmParams q = p;
    if (pass == 1) {
      q.bias = nullptr; q.accumulate = 0; q.act = GCNB_ACT_LINEAR; q.thresh = 0u;
      q.softmax = 1; q.logits = p0.logits;
    }
    for (int c0 = 0; c0 < K4; c0 += kMaxPassCols) {
      const int w = (K4 - c0) < kMaxPassCols ? (K4 - c0) : kMaxPassCols;
      q.col0 = c0; q.pcol0 = c0; q.nf4 = w / 4;
      const int nchunk = (q.nf4 + 31) / 32;
      if (pass == 0) {
        switch (nchunk) {
          case 1: launch_fixup<1>(ctx, q); break;
          case 2: launch_fixup<2>(ctx, q); break;
          case 3: launch_fixup<3>(ctx, q); break;
          default: launch_fixup<4>(ctx, q); break;
        }
      } else {
        switch (nchunk) {
          case 1: launch_softmax<1>(ctx, q, n_rows); break;
          case 2: launch_softmax<2>(ctx, q, n_rows); break;
          case 3: launch_softmax<3>(ctx, q, n_rows); break;
          default: launch_softmax<4>(ctx, q, n_rows); break;
        }
      }
      GCNB_LAUNCHED(ctx);
    }
  }
  return GCNB_OK;
}

}  // namespace

extern "C" size_t gcnb_spmm_workspace_bytes(const gcnb_csr* A, int32_t K) {
  if (!A) return 0;
  // partial-sum slots of the long rows: one pass of columns (engines 0/1) or the whole row (panel engine)
  return 256 + (size_t)A->n_slots * (((size_t)K + 3) / 4 * 4) * sizeof(float);
}

extern "C" int gcnb_spmm_engine_for(const gcnb_ctx* ctx, const gcnb_csr* A, int32_t ldb, int32_t K) {
  if (!ctx || !A) return GCNB_E_INVALID;
  int engine = 0, unroll = 0;
  pick_engine(ctx, A, ldb, ((K + 3) / 4) * 4, &engine, &unroll);
  return engine;
}

extern "C" int gcnb_csr_plan(const int32_t* rowptr, int32_t n_rows, int32_t chunk, int32_t* n_items,
                             int32_t* n_long, int32_t* n_slots, int32_t* items, int32_t* long_rows) {
  if (!rowptr || n_rows < 0 || chunk <= 0 || !n_items || !n_long || !n_slots) return GCNB_E_INVALID;
  long long ni = 0, nl = 0, ns = 0;
  for (int r = 0; r < n_rows; ++r) {
    const long long d = (long long)rowptr[r + 1] - rowptr[r];
    if (d < 0) return GCNB_E_INVALID;
    const long long pieces = d <= chunk ? 1 : (d + chunk - 1) / chunk;
    if (pieces > 1) {
      if (long_rows) {
        long_rows[3 * nl] = r;
        long_rows[3 * nl + 1] = (int32_t)ns;
        long_rows[3 * nl + 2] = (int32_t)pieces;
      }
      // even split keeps the pieces of one row the same size
      const long long per = (d + pieces - 1) / pieces;
      for (long long q = 0; q < pieces; ++q) {
        if (items) {
          const long long b = rowptr[r] + q * per;
          long long e = b + per;
          if (e > rowptr[r + 1]) e = rowptr[r + 1];
          items[4 * ni] = r; items[4 * ni + 1] = (int32_t)b; items[4 * ni + 2] = (int32_t)e;
          items[4 * ni + 3] = (int32_t)(ns + q);
        }
        ++ni;
      }
      ns += pieces;
      ++nl;
    } else {
      if (items) {
        items[4 * ni] = r; items[4 * ni + 1] = rowptr[r]; items[4 * ni + 2] = rowptr[r + 1];
        items[4 * ni + 3] = -1;
      }
      ++ni;
    }
    if (ni > 0x7fffffffLL || ns > 0x7fffffffLL) return GCNB_E_UNSUPPORTED;
  }
  *n_items = (int32_t)ni; *n_long = (int32_t)nl; *n_slots = (int32_t)ns;
  return GCNB_OK;
}

extern "C" int gcnb_spmm_csr_f32(gcnb_ctx* ctx, const gcnb_csr* A, const float* B, int32_t ldb, float* C,
                                 int32_t ldc, int32_t K, const gcnb_epilogue* epi) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, A && B && C, "null matrix");
  GCNB_REQUIRE(ctx, K > 0, "K must be positive");
  const int K4 = ((K + 3) / 4) * 4;
  GCNB_REQUIRE(ctx, (ldb % 4) == 0 && (ldc % 4) == 0 && ldb >= K4 && ldc >= K4, "ldb/ldc: multiple of 4, >= K rounded to 4");
  GCNB_REQUIRE(ctx, aligned16(B) && aligned16(C), "B and C must be 16-byte aligned");
  GCNB_REQUIRE(ctx, A->n_rows == 0 || (A->items && A->rowptr), "CSR not planned");
  GCNB_REQUIRE(ctx, A->nnz == 0 || (A->colidx && A->val), "CSR arrays missing");
  GCNB_REQUIRE(ctx, A->n_long == 0 || A->long_rows, "long row table missing");
  if (epi) {
    GCNB_REQUIRE(ctx, !epi->bias || aligned16(epi->bias), "bias must be 16-byte aligned");
    GCNB_REQUIRE(ctx, epi->dropout_p >= 0.f && epi->dropout_p < 1.f, "dropout_p in [0,1)");
    GCNB_REQUIRE(ctx, !epi->logits || aligned16(epi->logits), "logits must be 16-byte aligned");
  }
  if (A->n_rows == 0) return GCNB_OK;
  const size_t need = gcnb_spmm_workspace_bytes(A, K);
  int engine, unroll;
  pick_engine(ctx, A, ldb, K4, &engine, &unroll);
  if (A->n_slots > 0 || engine == 1) {
    if (!ctx->ws || ctx->ws_bytes < need)
      return gcnb_fail(ctx, GCNB_E_WORKSPACE, "spmm needs %s%lld workspace bytes, have %lld", "", (long long)need,
                       (long long)ctx->ws_bytes);
  }
  ProfScope scope(ctx, A->tag >= 0 && A->tag < GCNB_NTAGS ? A->tag : GCNB_TAG_SPMM_A);
  SpmmParams p;
  memset(&p, 0, sizeof(p));
  p.items = reinterpret_cast<const int4*>(A->items);
  p.n_items = A->n_items;
  p.col = A->colidx;
  p.val = A->val;
  p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.K = K;
  p.counter = reinterpret_cast<int*>(ctx->ws);
  p.partial = ctx->ws ? reinterpret_cast<float*>(reinterpret_cast<char*>(ctx->ws) + 256) : nullptr;
  p.long_rows = A->long_rows;
  p.n_long = A->n_long;
  if (epi) {
    p.bias = epi->bias; p.act = epi->act; p.softmax = epi->softmax; p.accumulate = epi->accumulate;
    if (epi->dropout_p > 0.f) {
      p.thresh = dropout_threshold(epi->dropout_p);
      if (p.thresh == 0u) p.thresh = 1u;
      p.scale = 1.f / (1.f - epi->dropout_p);
    }
    p.seed = epi->seed; p.row0 = epi->row0; p.logits = epi->logits;
  }
  if (!p.softmax) p.push = gcnb_take_push(ctx, K);  // a softmax output is never a convolution operand
  // a row softmax over more than one register pass of columns runs as a pass of its own after the product + bias
  const bool wide_softmax = p.softmax && K > kMaxPassCols;
  float* const wide_logits = p.logits;
  if (wide_softmax) {
    p.softmax = 0;
    p.logits = nullptr;
  }
  if (engine == 2) {
    const int rc = launch_panels(ctx, p, A->n_rows, K4, unroll, A->engine == -2);
    if (rc != GCNB_OK || !wide_softmax) return rc;
  }
  for (int c0 = 0; engine != 2 && c0 < K4; c0 += kMaxPassCols) {
    const int w = (K4 - c0) < kMaxPassCols ? (K4 - c0) : kMaxPassCols;
    p.col0 = c0;
    p.nf4 = w / 4;
    p.ldp = w;
    const int nchunk = (p.nf4 + 31) / 32;
    int rc;
    switch (nchunk) {
      case 1: rc = launch_pass<1>(ctx, p, engine, unroll); break;
      case 2: rc = launch_pass<2>(ctx, p, engine, unroll); break;
      case 3: rc = launch_pass<3>(ctx, p, engine, unroll); break;
      default: rc = launch_pass<4>(ctx, p, engine, unroll); break;
    }
    if (rc != GCNB_OK) return rc;
  }
  if (wide_softmax) {
    p.col0 = 0;
    p.k4 = K4;
    p.logits = wide_logits;
    row_softmax_wide_kernel<<<cdiv(A->n_rows, kWarpsPerCta), kWarpsPerCta * 32, 0, ctx->stream>>>(p, A->n_rows);
    GCNB_LAUNCHED(ctx);
  }
  return GCNB_OK;
}

// C[:, col0 : col0 + width] = epilogue(A . XP[:, 0 : width]) over all rows of the replicated A_hat; row i lands in rank
// (i / n_pad)'s copy of C (peer memory, csrc/peer.cu).  Always the panel engine: its lane groups own (row item, column
// panel) pairs, so a column slice is simply fewer panels, and every row is still summed in CSR order by one group.
extern "C" int gcnb_spmm_csr_sliced_f32(gcnb_ctx* ctx, const gcnb_csr* A, const float* XP, int32_t ldp, float* C,
                                        int32_t ldc, int32_t K, int32_t col0, int32_t width, int32_t n_pad,
                                        const gcnb_epilogue* epi) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, ctx->peer_world >= 2, "no peer arena attached (gcnb_peer_setup)");
  GCNB_REQUIRE(ctx, A && XP && C, "null matrix");
  GCNB_REQUIRE(ctx, K > 0 && col0 >= 0 && width >= 0 && n_pad > 0, "bad slice");
  GCNB_REQUIRE(ctx, col0 % 4 == 0 && width % 4 == 0 && ldp % 4 == 0 && ldc % 4 == 0 && ldp >= width && col0 + width <= ldc,
               "slice: columns in multiples of 4 inside C");
  GCNB_REQUIRE(ctx, aligned16(XP) && aligned16(C), "XP and C must be 16-byte aligned");
  GCNB_REQUIRE(ctx, (long long)n_pad * ctx->peer_world >= A->n_rows, "n_pad * world must cover the rows of A");
  GCNB_REQUIRE(ctx, A->n_rows == 0 || (A->items && A->rowptr), "CSR not planned");
  GCNB_REQUIRE(ctx, A->nnz == 0 || (A->colidx && A->val), "CSR arrays missing");
  GCNB_REQUIRE(ctx, A->n_long == 0 || A->long_rows, "long row table missing");
  if (epi) {
    GCNB_REQUIRE(ctx, !epi->softmax && epi->accumulate == 0 && epi->dropout_p == 0.f && !epi->logits,
                 "sliced product: bias + activation epilogue only");
    GCNB_REQUIRE(ctx, !epi->bias || aligned16(epi->bias), "bias must be 16-byte aligned");
  }
  if (A->n_rows == 0 || width == 0) return GCNB_OK;
  if (A->n_slots > 0) {
    const size_t need = gcnb_spmm_workspace_bytes(A, width);
    if (!ctx->ws || ctx->ws_bytes < need)
      return gcnb_fail(ctx, GCNB_E_WORKSPACE, "spmm needs %s%lld workspace bytes, have %lld", "", (long long)need,
                       (long long)ctx->ws_bytes);
  }
  ProfScope scope(ctx, A->tag >= 0 && A->tag < GCNB_NTAGS ? A->tag : GCNB_TAG_SPMM_A);
  SpmmParams p;
  memset(&p, 0, sizeof(p));
  p.items = reinterpret_cast<const int4*>(A->items);
  p.n_items = A->n_items;
  p.col = A->colidx;
  p.val = A->val;
  p.B = XP; p.ldb = ldp; p.C = C; p.ldc = ldc; p.K = K;
  p.counter = reinterpret_cast<int*>(ctx->ws);
  p.partial = ctx->ws ? reinterpret_cast<float*>(reinterpret_cast<char*>(ctx->ws) + 256) : nullptr;
  p.long_rows = A->long_rows;
  p.n_long = A->n_long;
  p.out_col0 = col0;
  p.n_pad = n_pad;
  const size_t span = ((size_t)n_pad - 1) * (size_t)ldc * sizeof(float) + (size_t)(col0 + width) * sizeof(float);
  for (int q = 0; q < ctx->peer_world; ++q) {
    p.peer_C[q] = reinterpret_cast<float*>(gcnb_peer_translate(ctx, C, q, span));
    GCNB_REQUIRE(ctx, p.peer_C[q] != nullptr, "C is not inside the peer arena");
  }
  if (epi) {
    p.bias = epi->bias; p.act = epi->act;
  }
  // Gather engine of the slice: the panel engine unless "spmm_sliced_engine" names another (0 / 1; both verified
  // bit-identical).  Measured at N = 2M, degree 64, slices of 64 columns on 8 GPUs (profiles/r2b_bench_c4_n8_*.json,
  // r2d_bench_c4_n8_slice_bulk.json): 32-column panels 27.2 ms per product (the 256 MB panel thrashes L2), 64-column
  // panels 13.7 ms, bulk-copy engine with 256-byte rows 15.6 ms -- against 4.7 ms + 6.6 ms of all-gather for the
  // row-block product, which is why the engine picks the all-gather exchange for graphs that large.
  int engine = ctx->spmm_sliced_engine;
  if (engine != 0 && engine != 1) engine = 2;
  if (engine == 2) {
    const bool big = (size_t)A->n_cols * 128 > ((size_t)80 << 20) && width >= 64 && ctx->spmm_panel == 32;
    const int saved = ctx->spmm_panel;
    if (big) ctx->spmm_panel = 64;  // panel no longer L2-resident: gather 256-byte rows
    const int rc = launch_panels(ctx, p, A->n_rows, width, ctx->spmm_unroll, false);
    ctx->spmm_panel = saved;
    return rc;
  }
  if (engine == 1 && (!ctx->ws || ctx->ws_bytes < 256))
    return gcnb_fail(ctx, GCNB_E_WORKSPACE, "spmm needs %s%lld workspace bytes, have %lld", "", 256LL, (long long)ctx->ws_bytes);
  for (int c0 = 0; c0 < width; c0 += kMaxPassCols) {
    const int w = (width - c0) < kMaxPassCols ? (width - c0) : kMaxPassCols;
    p.col0 = c0;
    p.pcol0 = 0;
    p.nf4 = w / 4;
    p.ldp = w;
    int rc;
    switch ((p.nf4 + 31) / 32) {
      case 1: rc = launch_pass<1>(ctx, p, engine, 2); break;
      case 2: rc = launch_pass<2>(ctx, p, engine, 2); break;
      case 3: rc = launch_pass<3>(ctx, p, engine, 2); break;
      default: rc = launch_pass<4>(ctx, p, engine, 2); break;
    }
    if (rc != GCNB_OK) return rc;
  }
  return GCNB_OK;
}

extern "C" int gcnb_row_softmax_f32(gcnb_ctx* ctx, float* C, int32_t ldc, int32_t n_rows, int32_t K, float* logits) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, C && K > 0 && n_rows >= 0, "bad argument");
  const int K4 = ((K + 3) / 4) * 4;
  GCNB_REQUIRE(ctx, ldc % 4 == 0 && ldc >= K4 && aligned16(C) && (!logits || aligned16(logits)), "ldc / alignment");
  if (n_rows == 0) return GCNB_OK;
  ProfScope scope(ctx, GCNB_TAG_LOSS);
  SpmmParams p;
  memset(&p, 0, sizeof(p));
  p.C = C; p.ldc = ldc; p.K = K; p.k4 = K4; p.nf4 = K4 / 4; p.softmax = 1; p.logits = logits;
  if (K4 > kMaxPassCols) {
    row_softmax_wide_kernel<<<cdiv(n_rows, kWarpsPerCta), kWarpsPerCta * 32, 0, ctx->stream>>>(p, n_rows);
  } else {
    switch ((p.nf4 + 31) / 32) {
      case 1: launch_softmax<1>(ctx, p, n_rows); break;
      case 2: launch_softmax<2>(ctx, p, n_rows); break;
      case 3: launch_softmax<3>(ctx, p, n_rows); break;
      default: launch_softmax<4>(ctx, p, n_rows); break;
    }
  }
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}
