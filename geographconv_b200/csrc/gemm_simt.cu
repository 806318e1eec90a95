// gemm_simt.cu -- fp32 CUDA-core GEMM with fused bias/activation and deterministic split-K.
//
// Replaces T.dot (reference gcnmodel.py:126,149,215) and the dgrad/wgrad products of its gradient
// for every shape and layout the tcgen05 path (gemm_tc.cu) is not instantiated for.  It is a
// device kernel, not a fallback to the host: exact fp32 FMA accumulation, any M/N/K, any
// combination of transposes.
//
// Tiling: 128 x 128 x 8 per CTA, 256 threads, 8 x 8 outputs per thread held as a 2 x 2 grid of
// 4 x 4 blocks (64 apart) so shared-memory reads are conflict-free LDS.128; register-staged double
// buffering of the global loads.  wgrad (K = number of graph nodes) is split over K: every split
// writes its own M x N partial to the workspace and a second kernel adds them in split order.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 8, NT = 256;

struct GemmParams {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  long long sa_m, sa_k;  // element strides of op(A)(m,k)
  long long sb_k, sb_n;  // element strides of op(B)(k,n)
  int ldc;
  int accumulate;
  const float* bias;
  int act;
  int k_per_split;  // multiple of BK
  float* partial;   // non-null: write raw tile sums to partial[z][M][N]
};

// AK: op(A) is K-contiguous (sa_k == 1); BN_: op(B) is N-contiguous (sb_n == 1)
template <bool AK, bool BNC>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = min(p.K, kbeg + p.k_per_split);
  const int tx = tid & 15, ty = tid >> 4;

  // global -> register staging: 4 elements of A and 4 of B per thread per k-tile
  int a_m[4], a_k[4], b_k[4], b_n[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (AK) { a_m[i] = tid >> 1; a_k[i] = (tid & 1) * 4 + i; }
    else    { a_m[i] = (tid & 31) * 4 + i; a_k[i] = tid >> 5; }
    if (BNC) { b_n[i] = (tid & 31) * 4 + i; b_k[i] = tid >> 5; }
    else     { b_n[i] = tid >> 1; b_k[i] = (tid & 1) * 4 + i; }
  }
  float ra[4], rb[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + a_m[i], k = k0 + a_k[i];
      ra[i] = (m < p.M && k < kend) ? __ldg(p.A + (long long)m * p.sa_m + (long long)k * p.sa_k) : 0.f;
      const int n = n0 + b_n[i], kb = k0 + b_k[i];
      rb[i] = (n < p.N && kb < kend) ? __ldg(p.B + (long long)kb * p.sb_k + (long long)n * p.sb_n) : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[buf][a_k[i]][a_m[i]] = ra[i];
      Bs[buf][b_k[i]][b_n[i]] = rb[i];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  int buf = 0;
  if (kbeg < kend) {
    gload(kbeg);
    sstore(0);
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4 + 64]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4 + 64]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 4 + (i & 3) + (i >> 2) * 64;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 4 + (j & 3) + (j >> 2) * 64;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.partial) {
        p.partial[((long long)blockIdx.z * p.M + m) * p.N + n] = v;
      } else {
        float* dst = p.C + (long long)m * p.ldc + n;
        if (p.accumulate) v += *dst;
        else {
          if (p.bias) v += __ldg(p.bias + n);
          v = act_apply(p.act, v);
        }
        *dst = v;
      }
    }
  }
}

__global__ void gemm_splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, float* C,
                                          int ldc, int accumulate, const float* __restrict__ bias, int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(long long)z * M * N + i];
  float* dst = C + (long long)m * ldc + n;
  if (accumulate) s += *dst;
  else {
    if (bias) s += bias[n];
    s = act_apply(act, s);
  }
  *dst = s;
}

// split-K factor; sized for the 148 SMs of a B200 (also used to size the workspace)
int pick_splits(int M, int N, int K) {
  constexpr int kSms = 148;
  const long long tiles = (long long)cdiv(M, BM) * cdiv(N, BN);
  if (tiles >= kSms || K < 4096) return 1;
  long long s = (2LL * kSms + tiles - 1) / tiles;
  const long long maxs = K / 512 > 0 ? K / 512 : 1;
  if (s > maxs) s = maxs;
  if (s > 512) s = 512;
  return (int)(s < 1 ? 1 : s);
}

}  // namespace

size_t gcnb_gemm_tc_workspace_bytes(int N, int K);
size_t gcnb_wgrad_tc_workspace_bytes(int M, int N, int K);

extern "C" size_t gcnb_gemm_workspace_bytes(int32_t transA, int32_t M, int32_t N, int32_t K) {
  const int s = pick_splits(M, N, K);
  const size_t simt = s <= 1 ? 0 : (size_t)s * M * N * sizeof(float);
  // tcgen05 paths: transposed weight copy (activation x weight) or split-K partial tiles (wgrad)
  const size_t tc = transA ? (M <= 4096 && N <= 4096 ? gcnb_wgrad_tc_workspace_bytes(M, N, K) : 0)
                           : gcnb_gemm_tc_workspace_bytes(N, K);
  return simt > tc ? simt : tc;
}

int gcnb_gemm_simt(gcnb_ctx* ctx, int transA, int transB, int M, int N, int K, const float* A, int lda,
                   const float* B, int ldb, float* C, int ldc, int accumulate, const float* bias, int act) {
  GemmParams p;
  p.A = A; p.B = B; p.C = C; p.M = M; p.N = N; p.K = K;
  p.sa_m = transA ? 1 : lda; p.sa_k = transA ? lda : 1;
  p.sb_k = transB ? 1 : ldb; p.sb_n = transB ? ldb : 1;
  p.ldc = ldc; p.accumulate = accumulate; p.bias = bias; p.act = act;
  p.partial = nullptr;
  int splits = pick_splits(M, N, K);
  int kps = cdiv(cdiv(K, splits), BK) * BK;
  splits = cdiv(K, kps);
  if (splits > 1) {
    const size_t need = (size_t)splits * M * N * sizeof(float);
    if (!ctx->ws || ctx->ws_bytes < need)
      return gcnb_fail(ctx, GCNB_E_WORKSPACE, "split-K gemm needs %s%lld workspace bytes, have %lld", "",
                       (long long)need, (long long)ctx->ws_bytes);
    p.partial = reinterpret_cast<float*>(ctx->ws);
  }
  p.k_per_split = splits > 1 ? kps : cdiv(K, BK) * BK;
  dim3 grid(cdiv(N, BN), cdiv(M, BM), splits);
  const bool AK = !transA, BNC = !transB;
  if (AK && BNC) gemm_simt_kernel<true, true><<<grid, NT, 0, ctx->stream>>>(p);
  else if (AK && !BNC) gemm_simt_kernel<true, false><<<grid, NT, 0, ctx->stream>>>(p);
  else if (!AK && BNC) gemm_simt_kernel<false, true><<<grid, NT, 0, ctx->stream>>>(p);
  else gemm_simt_kernel<false, false><<<grid, NT, 0, ctx->stream>>>(p);
  GCNB_LAUNCHED(ctx);
  if (splits > 1) {
    const long long n = (long long)M * N;
    gemm_splitk_reduce_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(p.partial, splits, M, N, C, ldc, accumulate,
                                                                      bias, act);
    GCNB_LAUNCHED(ctx);
  }
  return GCNB_OK;
}
