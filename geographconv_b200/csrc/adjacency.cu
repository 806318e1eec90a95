// adjacency.cu -- A_hat = D^-1/2 (Adj - diag + I) D^-1/2 built on the GPU from an undirected edge list.
//
// Replaces the host-side construction of the normalised adjacency in the reference
// (gcnmain.py:115-128: nx.adjacency_matrix -> setdiag(0) -> setdiag(1) -> row sums -> 1/sqrt ->
// D * adj * D -> astype(float32)); SURVEY.md 8f rank 1.  The graph is unweighted (no edge carries
// a 'w' attribute, data.py:56,61), so every stored entry of adj is 1 and the row sum is the
// number of distinct neighbours plus the self loop.
//
// HBM-bound integer work, no sort library:
//   1. count      raw degree of every node: 1 (self loop) + one per incident edge end (duplicates included)
//   2. scan       exclusive prefix sum -> bucket offsets
//   3. scatter    every edge end drops its neighbour id into the bucket of its node (atomic cursor; the order
//                 inside a bucket is arbitrary and is erased by the next step)
//   4. sort       each bucket is sorted in place: a warp per bucket in shared memory (<= 128 entries) or a CTA per
//                 bucket in shared memory (<= 4096) with an all-ascending ("flip") bitonic network; hubs of
//                 power-law graphs (> 4096 entries, up to most of the edge list at alpha = 1.5) go through an n-bit
//                 presence bitmap instead: the whole grid sets the bits of a batch of hub buckets (blockIdx.y = hub),
//                 then one CTA per hub enumerates the set bits in ascending order -- O(n/32 + m) instead of
//                 O(m log^2 m), and the result holds the DISTINCT neighbours only; distinct entries are counted.
//                 Count and scatter aggregate their atomics per warp (__match_any_sync), so that a hub that owns 60%
//                 of the edge ends costs one atomic per warp, not one per edge end.
//   5. scan       distinct counts -> rowptr of A_hat
//   6. fill       distinct neighbours are compacted into colidx (ascending inside a row) and
//                 val = float32( (d_i * 1.0) * d_j ) with d = 1/sqrt(double(row nnz)): the float64 arithmetic and
//                 the final rounding of the reference, so values are bit-identical to SciPy's.
// The result does not depend on the atomic order (sorting erases it): bit-reproducible.
#include <limits.h>

#include "common.cuh"

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr int kWarpRowMax = 128;  // bucket sizes a single warp sorts in shared memory
constexpr int kCtaRowMax = 4096;  // bucket sizes a CTA sorts in shared memory; larger buckets use the bitmap path
constexpr int kSortThreads = 512;
constexpr int kBigCtas = 128;     // CTAs of the big-bucket kernel (each owns one n-bit bitmap in the workspace)

struct AdjWork {  // carved out of the caller's workspace (all 256-byte aligned)
  int* cnt;       // n+1: raw bucket sizes, later reused for the distinct counts
  int* rawptr;    // n+1
  int* cursor;    // n: slot cursor during the scatter, afterwards the effective bucket length (efflen: the sorted
                  //    length, or -distinct for a hub bucket that was compacted through the bitmap)
  int* biglist;   // n: buckets of more than 128 entries
  int* hublist;   // n: of those, the ones of more than 4096 entries
  int* tmp;       // scan block sums
  int* flags;     // [0] out-of-range edge seen, [1] number of big buckets, [2] number of hubs
  double* dinv;   // n
  unsigned* bitmaps;  // kBigCtas x ceil(n / 32)
  int* raw;       // 2E + n
};

__host__ size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ size_t carve(AdjWork* w, char* base, long long n_edges, int n) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  };
  const size_t n1 = (size_t)n + 1;
  const size_t nblocks = (n1 + kScanTile - 1) / kScanTile + 1;
  int* cnt = (int*)take(n1 * 4);
  int* rawptr = (int*)take(n1 * 4);
  int* cursor = (int*)take(n1 * 4);
  int* biglist = (int*)take(n1 * 4);
  int* hublist = (int*)take(n1 * 4);
  int* tmp = (int*)take(nblocks * 4);
  int* flags = (int*)take(256);
  double* dinv = (double*)take(n1 * 8);
  unsigned* bitmaps = (unsigned*)take((size_t)kBigCtas * ((n1 + 31) / 32) * 4);
  int* raw = (int*)take(((size_t)2 * (size_t)n_edges + n1) * 4);
  if (w) *w = AdjWork{cnt, rawptr, cursor, biglist, hublist, tmp, flags, dinv, bitmaps, raw};
  return off;
}

// ------------------------------------------------------------------ exclusive scan (int32)
__global__ void __launch_bounds__(kScanThreads) scan_block_sums(const int* __restrict__ in, int n, int* __restrict__ sums) {
  __shared__ int wsum[kScanThreads / 32];
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < kScanThreads / 32; ++i) t += wsum[i];
    sums[blockIdx.x] = t;
  }
}

// one CTA: exclusive scan of the block sums in place, chunk by chunk with a carry
__global__ void __launch_bounds__(1024) scan_sums_inplace(int* sums, int nb) {
  __shared__ int wtot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < nb; c0 += 1024) {
    const int i = c0 + threadIdx.x;
    const int v = i < nb ? sums[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int t = wtot[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (threadIdx.x >= o) t += y;
      }
      wtot[threadIdx.x] = t;  // inclusive over warps
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int incl = x + (warp > 0 ? wtot[warp - 1] : 0);
    const int carry = carry_s;
    if (i < nb) sums[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply(const int* __restrict__ in, int n, const int* __restrict__ sums,
                                                           int* __restrict__ out) {
  __shared__ int wtot[kScanThreads / 32];
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = base + i < n ? in[base + i] : 0;
    s += v[i];
  }
  int x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = x;
  __syncthreads();
  int off = sums[blockIdx.x] + x - s;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += wtot[w];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = off;
    off += v[i];
  }
}

int exclusive_scan(gcnb_ctx* ctx, const int* in, int* out, int n, int* tmp) {
  const int nb = cdiv(n, kScanTile);
  scan_block_sums<<<nb, kScanThreads, 0, ctx->stream>>>(in, n, tmp);
  GCNB_LAUNCHED(ctx);
  scan_sums_inplace<<<1, 1024, 0, ctx->stream>>>(tmp, nb);
  GCNB_LAUNCHED(ctx);
  scan_apply<<<nb, kScanThreads, 0, ctx->stream>>>(in, n, tmp, out);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

// ------------------------------------------------------------------ count / scatter
__global__ void adj_init(int* cnt, int* cursor, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    cnt[i] = 1;  // the unit self loop (gcnmain.py:117-120)
    cursor[i] = 1;
  } else if (i == n) {
    cnt[i] = 0;
  }
}

// warp-aggregated counter updates: the lanes that hit the same node elect a leader that adds their count once.
// Called under divergence (invalid / self-loop edges skip it): the group is formed among the lanes that are here.
__device__ __forceinline__ void agg_add(int* cnt, int node) {
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, node);
  if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(cnt + node, __popc(peers));
}
// the same for slot allocation: returns this lane's slot in the bucket of `node`
__device__ __forceinline__ int agg_slot(int* cursor, int node) {
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, node);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(cursor + node, __popc(peers));
  base = __shfl_sync(peers, base, leader);
  return base + __popc(peers & ((1u << lane) - 1u));
}

__global__ void adj_count(const int* __restrict__ u, const int* __restrict__ v, long long n_edges, int n, int* cnt,
                          int* flags) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const int a = u[e], b = v[e];
    if (a < 0 || a >= n || b < 0 || b >= n) {
      flags[0] = 1;
      continue;
    }
    if (a == b) continue;  // setdiag(0): existing self loops are replaced by the unit one
    agg_add(cnt, a);
    agg_add(cnt, b);
  }
}

__global__ void adj_self(const int* __restrict__ rawptr, int* raw, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) raw[rawptr[i]] = i;
}

__global__ void adj_scatter(const int* __restrict__ u, const int* __restrict__ v, long long n_edges, int n,
                            const int* __restrict__ rawptr, int* cursor, int* raw) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += stride) {
    const int a = u[e], b = v[e];
    if (a < 0 || a >= n || b < 0 || b >= n || a == b) continue;
    raw[rawptr[a] + agg_slot(cursor, a)] = b;
    raw[rawptr[b] + agg_slot(cursor, b)] = a;
  }
}

// ------------------------------------------------------------------ per-bucket sort
// All-ascending bitonic network: every comparator leaves the smaller key at the lower index, so the
// network sorts any length m as if it were padded with +inf up to the next power of two (a comparator
// whose upper index is >= m would never swap and is skipped).
template <typename Sync>
__device__ __forceinline__ void bitonic_flip_sort(int* a, int m, int tid, int nthreads, Sync sync) {
  int P = 1;
  while (P < m) P <<= 1;
  for (int k = 2; k <= P; k <<= 1) {
    const int hk = k >> 1;
    for (int t = tid; t < (P >> 1); t += nthreads) {
      const int i = ((t & ~(hk - 1)) << 1) | (t & (hk - 1));
      const int l = i ^ (k - 1);
      if (l < m) {
        const int x = a[i], y = a[l];
        if (x > y) {
          a[i] = y;
          a[l] = x;
        }
      }
    }
    sync();
    for (int j = hk >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (P >> 1); t += nthreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i + j;
        if (l < m) {
          const int x = a[i], y = a[l];
          if (x > y) {
            a[i] = y;
            a[l] = x;
          }
        }
      }
      sync();
    }
  }
}

// distinct keys of a sorted bucket, counted by one warp
__device__ __forceinline__ int warp_count_distinct(const int* a, int m, int lane) {
  int cnt = 0;
  for (int i0 = 0; i0 < m; i0 += 32) {
    const int i = i0 + lane;
    bool f = false;
    if (i < m) f = (i == 0) || (a[i] != a[i - 1]);
    cnt += __popc(__ballot_sync(0xffffffffu, f));
  }
  return cnt;
}

__global__ void __launch_bounds__(256) adj_sort_small(const int* __restrict__ rawptr, int* raw, int n, int* distinct,
                                                      int* efflen, int* biglist, int* flags) {
  __shared__ int buf[8][kWarpRowMax];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * 8 + warp;
  if (r >= n) return;
  const int b = rawptr[r], m = rawptr[r + 1] - b;
  if (m > kWarpRowMax) {
    if (lane == 0) biglist[atomicAdd(flags + 1, 1)] = r;
    return;
  }
  int* a = buf[warp];
  for (int i = lane; i < m; i += 32) a[i] = raw[b + i];
  __syncwarp();
  bitonic_flip_sort(a, m, lane, 32, [] { __syncwarp(); });
  for (int i = lane; i < m; i += 32) raw[b + i] = a[i];
  const int d = warp_count_distinct(a, m, lane);
  if (lane == 0) {
    distinct[r] = d;
    efflen[r] = m;  // sorted, duplicates still in place
  }
}

// buckets of 129 .. 4096 entries: a CTA sorts them in shared memory; larger ones ("hubs") are queued for the bitmap path
__global__ void __launch_bounds__(kSortThreads) adj_sort_big(const int* __restrict__ rawptr, int* raw, int* distinct,
                                                             int* efflen, const int* __restrict__ biglist, int* flags,
                                                             int* hublist) {
  __shared__ int buf[kCtaRowMax];
  const int nbig = flags[1];
  const int tid = threadIdx.x;
  for (int q = blockIdx.x; q < nbig; q += gridDim.x) {
    const int r = biglist[q];
    const int b = rawptr[r], m = rawptr[r + 1] - b;
    int* a = raw + b;
    if (m > kCtaRowMax) {
      if (tid == 0) hublist[atomicAdd(flags + 2, 1)] = r;
      continue;
    }
    for (int i = tid; i < m; i += kSortThreads) buf[i] = a[i];
    __syncthreads();
    bitonic_flip_sort(buf, m, tid, kSortThreads, [] { __syncthreads(); });
    for (int i = tid; i < m; i += kSortThreads) a[i] = buf[i];
    __syncthreads();
    if (tid < 32) {
      const int d = warp_count_distinct(a, m, tid);
      if (tid == 0) {
        distinct[r] = d;
        efflen[r] = m;
      }
    }
    __syncthreads();
  }
}

// hubs, step 1: the whole grid sets presence bits.  blockIdx.y = hub of this batch (its own n-bit bitmap, zeroed before),
// blockIdx.x strides over the hub's raw bucket.  Neighbour ids spread over the bitmap words, so the RED.OR traffic of a
// 9M-entry bucket is spread over every SM and L2 slice instead of queueing behind one CTA.
__global__ void __launch_bounds__(256) adj_hub_setbits(const int* __restrict__ rawptr, const int* __restrict__ raw,
                                                       const int* __restrict__ hublist, const int* flags, int batch0,
                                                       size_t bm_words, unsigned* bitmaps) {
  const int h = batch0 + blockIdx.y;
  if (h >= flags[2]) return;
  const int r = hublist[h];
  const int b = rawptr[r], m = rawptr[r + 1] - b;
  unsigned* bm = bitmaps + (size_t)blockIdx.y * bm_words;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
    const int x = raw[b + i];
    atomicOr(bm + (x >> 5), 1u << (x & 31));
  }
}

// hubs, step 2: one CTA per hub turns the bitmap into the ascending list of DISTINCT neighbours at the head of the
// bucket (a thread owns a contiguous range of words: popcount, block scan, enumerate).  efflen = -d marks the bucket as
// compacted: adj_fill_hubs copies it with the whole grid, the warp-per-row fill skips it.
__global__ void __launch_bounds__(kSortThreads) adj_hub_enumerate(const int* __restrict__ rawptr, int* raw, int* distinct,
                                                                  int* efflen, const int* __restrict__ hublist,
                                                                  const int* flags, int batch0, int n, size_t bm_words,
                                                                  const unsigned* __restrict__ bitmaps) {
  __shared__ int wtot[kSortThreads / 32];
  const int h = batch0 + blockIdx.x;
  if (h >= flags[2]) return;
  const int r = hublist[h];
  int* a = raw + rawptr[r];
  const unsigned* bm = bitmaps + (size_t)blockIdx.x * bm_words;
  const int words = (n + 31) >> 5;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (words + kSortThreads - 1) / kSortThreads;
  const int w0 = min(tid * per, words), w1 = min(w0 + per, words);
  int c = 0;
  for (int w = w0; w < w1; ++w) c += __popc(bm[w]);
  int x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wtot[warp] = x;
  __syncthreads();
  int off = x - c, d = 0;
  for (int w = 0; w < kSortThreads / 32; ++w) {
    if (w < warp) off += wtot[w];
    d += wtot[w];
  }
  for (int w = w0; w < w1; ++w) {
    unsigned bits = bm[w];
    while (bits) {
      const int t = __ffs(bits) - 1;
      bits &= bits - 1;
      a[off++] = (w << 5) | t;
    }
  }
  if (tid == 0) {
    distinct[r] = d;
    efflen[r] = -d;
  }
}

// hubs, step 3 (after rowptr is known): entries are distinct and ascending, so the fill is a parallel copy + scale
__global__ void __launch_bounds__(256) adj_fill_hubs(const int* __restrict__ rawptr, const int* __restrict__ raw,
                                                     const int* __restrict__ rowptr, const double* __restrict__ dinv,
                                                     const int* __restrict__ hublist, const int* flags,
                                                     int* __restrict__ colidx, float* __restrict__ val) {
  const int h = blockIdx.y;
  if (h >= flags[2]) return;
  const int r = hublist[h];
  const int* a = raw + rawptr[r];
  const int out = rowptr[r], d = rowptr[r + 1] - out;
  const double di = dinv[r];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d; i += gridDim.x * blockDim.x) {
    const int x = a[i];
    colidx[out + i] = x;
    val[out + i] = (float)((di * 1.0) * dinv[x]);
  }
}

__global__ void adj_dinv(const int* __restrict__ rowptr, double* dinv, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int d = rowptr[i + 1] - rowptr[i];
    dinv[i] = d > 0 ? 1.0 / sqrt((double)d) : 0.0;  // inf -> 0 (gcnmain.py:123-125)
  }
}

// warp per row: compact the distinct neighbours, scale
__global__ void __launch_bounds__(256) adj_fill(const int* __restrict__ rawptr, const int* __restrict__ raw,
                                                const int* __restrict__ efflen, const int* __restrict__ rowptr,
                                                const double* __restrict__ dinv, int n,
                                                int* __restrict__ colidx, float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const int b = rawptr[r], m = efflen[r];
  if (m < 0) return;  // compacted hub: adj_fill_hubs
  const int* a = raw + b;
  const double di = dinv[r];
  int out = rowptr[r];
  for (int i0 = 0; i0 < m; i0 += 32) {
    const int i = i0 + lane;
    int x = 0;
    bool f = false;
    if (i < m) {
      x = a[i];
      f = (i == 0) || (x != a[i - 1]);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, f);
    if (f) {
      const int pos = out + __popc(mask & ((1u << lane) - 1u));
      colidx[pos] = x;
      val[pos] = (float)((di * 1.0) * dinv[x]);  // (D * adj) * D in float64, then astype(float32)
    }
    out += __popc(mask);
  }
}

}  // namespace

extern "C" size_t gcnb_adj_workspace_bytes(int64_t n_edges, int32_t n_nodes) {
  if (n_edges < 0 || n_nodes < 0) return 0;
  return carve(nullptr, nullptr, n_edges, n_nodes);
}

extern "C" int gcnb_adj_build_rows(gcnb_ctx* ctx, const int32_t* u, const int32_t* v, int64_t n_edges, int32_t n_nodes,
                                   void* work, size_t work_bytes, int32_t* rowptr, int64_t* nnz_host) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, n_edges >= 0 && n_nodes >= 0, "negative size");
  GCNB_REQUIRE(ctx, n_edges == 0 || (u && v), "null edge arrays");
  GCNB_REQUIRE(ctx, rowptr && nnz_host, "null output");
  GCNB_REQUIRE(ctx, 2 * (long long)n_edges + n_nodes < 0x7fffffffLL, "2*edges + nodes must fit int32");
  GCNB_REQUIRE(ctx, work && (reinterpret_cast<uintptr_t>(work) & 255) == 0, "workspace must be 256-byte aligned");
  if (work_bytes < carve(nullptr, nullptr, n_edges, n_nodes))
    return gcnb_fail(ctx, GCNB_E_WORKSPACE, "adjacency build needs %s%lld workspace bytes, have %lld", "",
                     (long long)carve(nullptr, nullptr, n_edges, n_nodes), (long long)work_bytes);
  ProfScope scope(ctx, GCNB_TAG_ELEM);
  AdjWork w;
  carve(&w, reinterpret_cast<char*>(work), n_edges, n_nodes);
  const int n = n_nodes;
  GCNB_CUDA(ctx, cudaMemsetAsync(w.flags, 0, 256, ctx->stream));
  adj_init<<<cdiv((long long)n + 1, 256), 256, 0, ctx->stream>>>(w.cnt, w.cursor, n);
  GCNB_LAUNCHED(ctx);
  const int egrid = n_edges > 0 ? (int)(((n_edges + 255) / 256 < 148 * 16) ? (n_edges + 255) / 256 : 148 * 16) : 0;
  if (egrid > 0) {
    adj_count<<<egrid, 256, 0, ctx->stream>>>(u, v, n_edges, n, w.cnt, w.flags);
    GCNB_LAUNCHED(ctx);
  }
  int rc = exclusive_scan(ctx, w.cnt, w.rawptr, n + 1, w.tmp);
  if (rc != GCNB_OK) return rc;
  if (n > 0) {
    adj_self<<<cdiv(n, 256), 256, 0, ctx->stream>>>(w.rawptr, w.raw, n);
    GCNB_LAUNCHED(ctx);
  }
  if (egrid > 0) {
    adj_scatter<<<egrid, 256, 0, ctx->stream>>>(u, v, n_edges, n, w.rawptr, w.cursor, w.raw);
    GCNB_LAUNCHED(ctx);
  }
  if (n > 0) {
    // distinct counts overwrite the raw counts (the bucket offsets are already in rawptr)
    adj_sort_small<<<cdiv(n, 8), 256, 0, ctx->stream>>>(w.rawptr, w.raw, n, w.cnt, w.cursor, w.biglist, w.flags);
    GCNB_LAUNCHED(ctx);
    adj_sort_big<<<kBigCtas, kSortThreads, 0, ctx->stream>>>(w.rawptr, w.raw, w.cnt, w.cursor, w.biglist, w.flags,
                                                             w.hublist);
    GCNB_LAUNCHED(ctx);
    // hubs (buckets of more than 4096 entries; none on a uniform graph): the count decides how many bitmap batches run
    int n_hubs = 0;
    GCNB_CUDA(ctx, cudaMemcpyAsync(&n_hubs, w.flags + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GCNB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t bm_words = ((size_t)n + 1 + 31) / 32;
    for (int b0 = 0; b0 < n_hubs; b0 += kBigCtas) {
      const int nb = n_hubs - b0 < kBigCtas ? n_hubs - b0 : kBigCtas;
      GCNB_CUDA(ctx, cudaMemsetAsync(w.bitmaps, 0, (size_t)nb * bm_words * 4, ctx->stream));
      adj_hub_setbits<<<dim3((unsigned)(ctx->sm_count * 2), (unsigned)nb), 256, 0, ctx->stream>>>(
          w.rawptr, w.raw, w.hublist, w.flags, b0, bm_words, w.bitmaps);
      GCNB_LAUNCHED(ctx);
      adj_hub_enumerate<<<nb, kSortThreads, 0, ctx->stream>>>(w.rawptr, w.raw, w.cnt, w.cursor, w.hublist, w.flags, b0, n,
                                                              bm_words, w.bitmaps);
      GCNB_LAUNCHED(ctx);
    }
  }
  rc = exclusive_scan(ctx, w.cnt, rowptr, n + 1, w.tmp);
  if (rc != GCNB_OK) return rc;
  if (n > 0) {
    adj_dinv<<<cdiv(n, 256), 256, 0, ctx->stream>>>(rowptr, w.dinv, n);
    GCNB_LAUNCHED(ctx);
  }
  int h_flags[2] = {0, 0};
  int h_nnz = 0;
  GCNB_CUDA(ctx, cudaMemcpyAsync(h_flags, w.flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream));
  GCNB_CUDA(ctx, cudaMemcpyAsync(&h_nnz, rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GCNB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_flags[0]) return gcnb_fail(ctx, GCNB_E_INVALID, "edge list holds a node id outside [0, %s%lld)", "", (long long)n);
  *nnz_host = h_nnz;
  return GCNB_OK;
}

extern "C" int gcnb_adj_fill_f32(gcnb_ctx* ctx, int64_t n_edges, int32_t n_nodes, const void* work,
                                 const int32_t* rowptr, int32_t* colidx, float* val) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, n_edges >= 0 && n_nodes >= 0, "negative size");
  GCNB_REQUIRE(ctx, work && rowptr, "null workspace / rowptr");
  if (n_nodes == 0) return GCNB_OK;
  GCNB_REQUIRE(ctx, colidx && val, "null output");
  ProfScope scope(ctx, GCNB_TAG_ELEM);
  AdjWork w;
  carve(&w, const_cast<char*>(reinterpret_cast<const char*>(work)), n_edges, n_nodes);
  adj_fill<<<cdiv(n_nodes, 8), 256, 0, ctx->stream>>>(w.rawptr, w.raw, w.cursor, rowptr, w.dinv, n_nodes, colidx, val);
  GCNB_LAUNCHED(ctx);
  int n_hubs = 0;
  GCNB_CUDA(ctx, cudaMemcpyAsync(&n_hubs, w.flags + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GCNB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_hubs > 0) {
    adj_fill_hubs<<<dim3(64u, (unsigned)n_hubs), 256, 0, ctx->stream>>>(w.rawptr, w.raw, rowptr, w.dinv, w.hublist, w.flags,
                                                                          colidx, val);
    GCNB_LAUNCHED(ctx);
  }
  return GCNB_OK;
}

// ---------------------------------------------------------------------------------------
// weighted graphs: nx.adjacency_matrix(..., weight='w') (gcnmain.py:115) yields edge weights when edges carry a
// 'w' attribute.  The caller hands the symmetric weighted CSR with its unit diagonal already in place
// (setdiag(0); setdiag(1), gcnmain.py:117-120 -- structure edits); the row sums, 1/sqrt and the two-sided scaling
// (gcnmain.py:122-128) run here in the reference's float64 with one final rounding to float32.
// ---------------------------------------------------------------------------------------
namespace {
// one thread per row: the float64 sum runs over the row in stored order, like SciPy's csr_matvec behind adj.sum(axis=1)
__global__ void adjw_dinv_kernel(const int* __restrict__ rowptr, const double* __restrict__ w, int n, double* dinv) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double s = 0.0;
  for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) s = __dadd_rn(s, w[k]);
  const double d = 1.0 / sqrt(s);
  dinv[r] = isinf(d) ? 0.0 : d;  // diags_sqrt[isinf] = 0 (gcnmain.py:124)
}
__global__ void adjw_scale_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ w,
                                  const double* __restrict__ dinv, int n, float* val) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const double di = dinv[r];
  for (int k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32)
    val[k] = (float)__dmul_rn(__dmul_rn(di, w[k]), dinv[col[k]]);  // (D * adj) * D, then astype(float32)
}
}  // namespace

extern "C" int gcnb_adj_normalize_weighted_f64(gcnb_ctx* ctx, const int32_t* rowptr, const int32_t* colidx,
                                               const double* weights, int32_t n_nodes, double* dinv_work, float* val) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, n_nodes >= 0, "negative size");
  if (n_nodes == 0) return GCNB_OK;
  GCNB_REQUIRE(ctx, rowptr && colidx && weights && dinv_work && val, "null pointer");
  ProfScope scope(ctx, GCNB_TAG_ELEM);
  adjw_dinv_kernel<<<cdiv(n_nodes, 256), 256, 0, ctx->stream>>>(rowptr, weights, n_nodes, dinv_work);
  GCNB_LAUNCHED(ctx);
  adjw_scale_kernel<<<cdiv(n_nodes, 8), 256, 0, ctx->stream>>>(rowptr, colidx, weights, dinv_work, n_nodes, val);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}
