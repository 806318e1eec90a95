// elementwise.cu -- the HBM-bound row-local kernels of the GCN step: highway gate mix and its
// backward, activation/dropout backward, bias-gradient column sums, softmax cross-entropy metrics
// and gradient, prediction gather/argmax, L1+L2 regularisation, Adam, dropout-mask dump.
// Each replaces a fused theano Elemwise / Softmax / AdvancedSubtensor1 / MaxAndArgmax node of the
// reference graph (gcnmodel.py:132-136,266,357,374-389,407); all are float4-vectorised streaming
// kernels whose roofline is HBM bandwidth.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float4 ld4(const float* base, size_t row, int ld, int c4) {
  return *reinterpret_cast<const float4*>(base + row * (size_t)ld + 4 * (size_t)c4);
}
__device__ __forceinline__ void st4(float* base, size_t row, int ld, int c4, float4 v) {
  *reinterpret_cast<float4*>(base + row * (size_t)ld + 4 * (size_t)c4) = v;
}

// Y = T*H + (1-T)*X            (MultiplicativeGatingLayer, gcnmodel.py:266)
__global__ void highway_mix_kernel(int n_rows, int nf4, const float* H, int ldh, const float* T, int ldt,
                                   const float* X, int ldx, float* Y, int ldy) {
  const long long total = (long long)n_rows * nf4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const size_t r = (size_t)(i / nf4);
    const int c = (int)(i % nf4);
    const float4 h = ld4(H, r, ldh, c), t = ld4(T, r, ldt, c), x = ld4(X, r, ldx, c);
    float4 y;
    y.x = t.x * h.x + (1.0f - t.x) * x.x;
    y.y = t.y * h.y + (1.0f - t.y) * x.y;
    y.z = t.z * h.z + (1.0f - t.z) * x.z;
    y.w = t.w * h.w + (1.0f - t.w) * x.w;
    st4(Y, r, ldy, c, y);
  }
}

__global__ void highway_bwd_kernel(int n_rows, int nf4, int ld, const float* dY, const float* X, const float* H,
                                   const float* T, int act, float* dH, float* dT, float* dX) {
  const long long total = (long long)n_rows * nf4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const size_t r = (size_t)(i / nf4);
    const int c = (int)(i % nf4);
    const float4 g = ld4(dY, r, ld, c), x = ld4(X, r, ld, c), h = ld4(H, r, ld, c), t = ld4(T, r, ld, c);
    float4 dh, dt, dx;
#define GCNB_HW_BWD(e)                                         \
  dh.e = g.e * t.e * act_grad_from_out(act, h.e);              \
  dt.e = g.e * (h.e - x.e) * t.e * (1.0f - t.e);               \
  dx.e = g.e * (1.0f - t.e);
    GCNB_HW_BWD(x) GCNB_HW_BWD(y) GCNB_HW_BWD(z) GCNB_HW_BWD(w)
#undef GCNB_HW_BWD
    st4(dH, r, ld, c, dh);
    st4(dT, r, ld, c, dt);
    st4(dX, r, ld, c, dx);
  }
}

// dZ = dY * keep*scale * act'(a), a recovered from the stored dropout(act(z))
__global__ void act_bwd_kernel(int n_rows, int nf4, int ld, const float* dY, const float* Yact, int act,
                               uint32_t thresh, float scale, uint64_t seed, int64_t row0, float* dZ) {
  const long long total = (long long)n_rows * nf4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const size_t r = (size_t)(i / nf4);
    const int c = (int)(i % nf4);
    const float4 g = ld4(dY, r, ld, c), y = ld4(Yact, r, ld, c);
    float4 o;
    if (thresh != 0u) {
      const uint4 d = dropout_draw(seed, row0 + (int64_t)r, (uint32_t)c);
      const float inv = 1.0f / scale;
      o.x = d.x < thresh ? g.x * scale * act_grad_from_out(act, y.x * inv) : 0.f;
      o.y = d.y < thresh ? g.y * scale * act_grad_from_out(act, y.y * inv) : 0.f;
      o.z = d.z < thresh ? g.z * scale * act_grad_from_out(act, y.z * inv) : 0.f;
      o.w = d.w < thresh ? g.w * scale * act_grad_from_out(act, y.w * inv) : 0.f;
    } else {
      o.x = g.x * act_grad_from_out(act, y.x);
      o.y = g.y * act_grad_from_out(act, y.y);
      o.z = g.z * act_grad_from_out(act, y.z);
      o.w = g.w * act_grad_from_out(act, y.w);
    }
    st4(dZ, r, ld, c, o);
  }
}

// column sums, stage 1: block b sums rows b, b+gridDim.x, ... ; 8 warps x 32 lanes, lanes over float4 columns
__global__ void __launch_bounds__(kThreads) colsum_stage1_kernel(int n_rows, int nf4, const float* A, int lda,
                                                                 float* partial) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < nf4; c0 += 32) {
    const int c = c0 + lane;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < nf4) {
      for (long long r = (long long)blockIdx.x * 8 + warp; r < n_rows; r += (long long)gridDim.x * 8) {
        const float4 v = ld4(A, (size_t)r, lda, c);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    red[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && c < nf4) {
      float4 t = red[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) { t.x += red[w][lane].x; t.y += red[w][lane].y; t.z += red[w][lane].z; t.w += red[w][lane].w; }
      reinterpret_cast<float4*>(partial + (size_t)blockIdx.x * nf4 * 4)[c] = t;
    }
    __syncthreads();
  }
}
// stage 2: out[c] (+)= sum over blocks of partial[b * stride + c].  A CTA owns 32 consecutive columns: warp w adds
// blocks w, w+8, ... (coalesced 128-byte reads), then the 8 warp sums are added in warp order: a fixed order, so the
// result is deterministic.
__global__ void __launch_bounds__(kThreads) colsum_stage2_kernel(int nblocks, int k, size_t stride, const float* partial,
                                                                 float* out, int accumulate) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (c < k)
    for (int b = warp; b < nblocks; b += 8) s += partial[(size_t)b * stride + c];
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && c < k) {
    float t = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red[w][lane];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// block-level column sum of per-thread float4 partials (8 warps x 32 lanes): warp 0 adds the 8 warps in order
__device__ __forceinline__ void block_colsum_store(float4 (*red)[32], float4 s, int warp, int lane, bool on, float4* dst) {
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && on) {
    float4 t = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) { t.x += red[w][lane].x; t.y += red[w][lane].y; t.z += red[w][lane].z; t.w += red[w][lane].w; }
    *dst = t;
  }
  __syncthreads();
}

// highway backward fused with the two bias gradients: a warp owns whole rows (lanes over float4 columns), every
// block keeps running column sums of dHpre and dTpre and writes them to partial[block][2][nf4*4]; stage 2 adds the
// blocks in order (deterministic).  Saves re-reading dHpre and dTpre for gcnb_colsum_f32.
template <int NCH>
__global__ void __launch_bounds__(kThreads) highway_bwd_colsum_kernel(int n_rows, int nf4, int ld, const float* dY,
                                                                      const float* X, const float* H, const float* T,
                                                                      int act, float* dH, float* dT, float* dX,
                                                                      float* partial, const PushPlan pp) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 sH[NCH], sT[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) sH[ch] = sT[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = (long long)blockIdx.x * 8 + warp; r < n_rows; r += (long long)gridDim.x * 8) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int c = lane + 32 * ch;
      if (c < nf4) {
        const float4 g = ld4(dY, (size_t)r, ld, c), x = ld4(X, (size_t)r, ld, c), h = ld4(H, (size_t)r, ld, c),
                     t = ld4(T, (size_t)r, ld, c);
        float4 dh, dt, dx;
#define GCNB_HW_BWD(e)                                         \
  dh.e = g.e * t.e * act_grad_from_out(act, h.e);              \
  dt.e = g.e * (h.e - x.e) * t.e * (1.0f - t.e);               \
  dx.e = g.e * (1.0f - t.e);                                   \
  sH[ch].e += dh.e;                                            \
  sT[ch].e += dt.e;
        GCNB_HW_BWD(x) GCNB_HW_BWD(y) GCNB_HW_BWD(z) GCNB_HW_BWD(w)
#undef GCNB_HW_BWD
        // dHpre only feeds the graph convolution V = A^T.dHpre: in a feature-sliced run it goes straight to the
        // ranks that own its columns (NVLink stores riding under this kernel's HBM traffic), not to local memory
        if (pp.on) push_store_f4(pp, r, 4 * c, dh);
        else st4(dH, (size_t)r, ld, c, dh);
        st4(dT, (size_t)r, ld, c, dt);
        st4(dX, (size_t)r, ld, c, dx);
      }
    }
  }
  float4* p4 = reinterpret_cast<float4*>(partial) + (size_t)blockIdx.x * 2 * nf4;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int c = lane + 32 * ch;
    block_colsum_store(red, sH[ch], warp, lane, c < nf4, p4 + c);
    block_colsum_store(red, sT[ch], warp, lane, c < nf4, p4 + nf4 + c);
  }
}

// activation / dropout backward fused with the bias gradient (column sums of dZ), same scheme
template <int NCH>
__global__ void __launch_bounds__(kThreads) act_bwd_colsum_kernel(int n_rows, int nf4, int ld, const float* dY,
                                                                  const float* Yact, int act, uint32_t thresh, float scale,
                                                                  uint64_t seed, int64_t row0, float* dZ, float* partial,
                                                                  const PushPlan pp) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 sZ[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) sZ[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float inv = 1.0f / scale;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < n_rows; r += (long long)gridDim.x * 8) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int c = lane + 32 * ch;
      if (c < nf4) {
        const float4 g = ld4(dY, (size_t)r, ld, c), y = ld4(Yact, (size_t)r, ld, c);
        float4 o;
        if (thresh != 0u) {
          const uint4 d = dropout_draw(seed, row0 + (int64_t)r, (uint32_t)c);
          o.x = d.x < thresh ? g.x * scale * act_grad_from_out(act, y.x * inv) : 0.f;
          o.y = d.y < thresh ? g.y * scale * act_grad_from_out(act, y.y * inv) : 0.f;
          o.z = d.z < thresh ? g.z * scale * act_grad_from_out(act, y.z * inv) : 0.f;
          o.w = d.w < thresh ? g.w * scale * act_grad_from_out(act, y.w * inv) : 0.f;
        } else {
          o.x = g.x * act_grad_from_out(act, y.x);
          o.y = g.y * act_grad_from_out(act, y.y);
          o.z = g.z * act_grad_from_out(act, y.z);
          o.w = g.w * act_grad_from_out(act, y.w);
        }
        sZ[ch].x += o.x; sZ[ch].y += o.y; sZ[ch].z += o.z; sZ[ch].w += o.w;
        if (pp.on) push_store_f4(pp, r, 4 * c, o);  // operand of the next graph convolution: to its column owners
        else st4(dZ, (size_t)r, ld, c, o);
      }
    }
  }
  float4* p4 = reinterpret_cast<float4*>(partial) + (size_t)blockIdx.x * nf4;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int c = lane + 32 * ch;
    block_colsum_store(red, sZ[ch], warp, lane, c < nf4, p4 + c);
  }
}

// cross-entropy metrics: one warp per gathered row
__global__ void __launch_bounds__(kThreads) xent_metrics_stage1_kernel(const float* P, int ldp, int C, const int* idx,
                                                                       const int* labels, int n_idx, float* partial) {
  __shared__ float red[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float loss = 0.f, correct = 0.f;
  for (long long i = (long long)blockIdx.x * 8 + warp; i < n_idx; i += (long long)gridDim.x * 8) {
    const float* row = P + (size_t)idx[i] * ldp;
    const int y = labels[i];
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v = row[c];
      if (v > best) { best = v; arg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) {
      loss += -logf(row[y]);
      correct += (arg == y) ? 1.f : 0.f;
    }
  }
  if (lane == 0) { red[0][warp] = loss; red[1][warp] = correct; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, c = 0.f;
    for (int w = 0; w < 8; ++w) { l += red[0][w]; c += red[1][w]; }
    partial[2 * blockIdx.x] = l;
    partial[2 * blockIdx.x + 1] = c;
  }
}
__global__ void xent_metrics_stage2_kernel(int nblocks, const float* partial, float* metrics) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  // double accumulation of the per-block sums keeps the reported mean stable for 10^5..10^6 rows; one warp, lane l
  // adds blocks l, l+32, ..., then a fixed-order shuffle tree: deterministic, and 30x shorter than one thread walking
  // all blocks (60 us, visible at 6 ms per step on 8 GPUs)
  double l = 0.0, c = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 32) { l += partial[2 * b]; c += partial[2 * b + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l += __shfl_down_sync(0xffffffffu, l, o);
    c += __shfl_down_sync(0xffffffffu, c, o);
  }
  if (threadIdx.x == 0) {
    metrics[0] += (float)l;
    metrics[1] += (float)c;
  }
}

__global__ void __launch_bounds__(kThreads) xent_grad_kernel(const float* P, int ldp, int C, const int* idx,
                                                             const int* labels, int n_idx, float inv_n, float* G,
                                                             int ldg) {
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n_idx) return;
  const int r = idx[i], y = labels[i];
  const float* row = P + (size_t)r * ldp;
  float* g = G + (size_t)r * ldg;
  for (int c = lane; c < C; c += 32) atomicAdd(g + c, (row[c] - (c == y ? 1.f : 0.f)) * inv_n);
}

// dense form of the loss gradient: row_label[r] = class of local row r if it is a training row, else -1.  A warp per
// row writes the whole row of G (zeros for non-training rows), locally for the bias gradient and -- armed -- into the
// panel buffers of the column owners for the graph convolution A^T.G that follows.
__global__ void __launch_bounds__(kThreads) xent_grad_dense_kernel(const float* P, int ldp, int C, int n_rows,
                                                                   const int* row_label, float inv_n, float* G, int ldg,
                                                                   const PushPlan pp) {
  const int lane = threadIdx.x & 31;
  const int nf4 = (C + 3) >> 2;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < n_rows; r += (long long)gridDim.x * 8) {
    const int y = row_label[r];
    for (int c4 = lane; c4 < nf4; c4 += 32) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0) {
        const float4 pr = ld4(P, (size_t)r, ldp, c4);
        const int c = 4 * c4;
        g.x = c + 0 < C ? (pr.x - (c + 0 == y ? 1.f : 0.f)) * inv_n : 0.f;
        g.y = c + 1 < C ? (pr.y - (c + 1 == y ? 1.f : 0.f)) * inv_n : 0.f;
        g.z = c + 2 < C ? (pr.z - (c + 2 == y ? 1.f : 0.f)) * inv_n : 0.f;
        g.w = c + 3 < C ? (pr.w - (c + 3 == y ? 1.f : 0.f)) * inv_n : 0.f;
      }
      st4(G, (size_t)r, ldg, c4, g);
      if (pp.on) push_store_f4(pp, r, 4 * c4, g);
    }
  }
}

__global__ void __launch_bounds__(kThreads) gather_argmax_kernel(const float* P, int ldp, int C, const int* idx,
                                                                 int n_idx, long long* preds, float* probs) {
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n_idx) return;
  const float* row = P + (size_t)idx[i] * ldp;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float v = row[c];
    if (probs) probs[(size_t)i * C + c] = v;
    if (v > best) { best = v; arg = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  if (lane == 0) preds[i] = arg == 0x7fffffff ? 0 : arg;
}

__global__ void __launch_bounds__(kThreads) l1l2_kernel(const float* W, float* G, long long n, float coef,
                                                        float* reg_sum) {
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float w = W[i];
    s += fabsf(w) + w * w;
    const float sg = w > 0.f ? 1.f : (w < 0.f ? -1.f : 0.f);
    G[i] += coef * (sg + 2.f * w);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(reg_sum, t);
  }
}

// lasagne.updates.adam: t <- t+1; a_t = lr*sqrt(1-b2^t)/(1-b1^t)
__global__ void adam_tick_kernel(float* state, float lr, float b1, float b2) {
  const float t = state[0] + 1.0f;
  state[0] = t;
  state[1] = lr * sqrtf(1.0f - powf(b2, t)) / (1.0f - powf(b1, t));
}
__global__ void __launch_bounds__(kThreads) adam_kernel(float4* P, const float4* G, float4* M, float4* V,
                                                        long long n4, const float* state, float b1, float b2,
                                                        float eps) {
  const float a_t = state[1];
  const float c1 = 1.0f - b1, c2 = 1.0f - b2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g = G[i];
    float4 m = M[i], v = V[i], p = P[i];
#define GCNB_ADAM(e)                       \
  m.e = b1 * m.e + c1 * g.e;               \
  v.e = b2 * v.e + c2 * g.e * g.e;         \
  p.e = p.e - a_t * m.e / (sqrtf(v.e) + eps);
    GCNB_ADAM(x) GCNB_ADAM(y) GCNB_ADAM(z) GCNB_ADAM(w)
#undef GCNB_ADAM
    M[i] = m; V[i] = v; P[i] = p;
  }
}

__global__ void dropout_mask_kernel(int n_rows, int k, uint32_t thresh, uint64_t seed, int64_t row0, uint8_t* mask) {
  const int ngrp = (k + 3) / 4;
  const long long total = (long long)n_rows * ngrp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ngrp;
    const int g = (int)(i % ngrp);
    const uint4 d = dropout_draw(seed, row0 + r, (uint32_t)g);
    const uint32_t dv[4] = {d.x, d.y, d.z, d.w};
    for (int e = 0; e < 4; ++e)
      if (4 * g + e < k) mask[r * k + 4 * g + e] = (thresh == 0u || dv[e] < thresh) ? 1 : 0;
  }
}

// dst[i] = src[idx[i]] (gather) or dst[idx[i]] = src[i] (scatter), float4 rows
// dense[r, col[k]] = val[k] for the nonzeros k of CSR row r (warp per row; dense was cleared beforehand)
__global__ void __launch_bounds__(kThreads) csr_to_dense_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                                const float* __restrict__ val, int n_rows, float* dense,
                                                                int ld) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  float* out = dense + (size_t)r * ld;
  for (int k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) out[col[k]] = val[k];
}

__global__ void move_rows_kernel(const float* src, int ld_src, const int* idx, int n_idx, int nf4, float* dst,
                                 int ld_dst, int scatter) {
  const long long total = (long long)n_idx * nf4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / nf4), c = (int)(t % nf4);
    const int r = idx[i];
    const size_t rs = scatter ? (size_t)i : (size_t)r, rd = scatter ? (size_t)r : (size_t)i;
    st4(dst, rd, ld_dst, c, ld4(src, rs, ld_src, c));
  }
}

// column ids that fit 16 bits cross PCIe as uint16 and are widened here (8 per thread: one 128-bit load, two stores)
__global__ void __launch_bounds__(kThreads) expand_u16_kernel(const uint16_t* __restrict__ src, long long n, int* dst) {
  const long long n8 = n >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(src)[i];
    int4 lo = make_int4((int)(v.x & 0xffffu), (int)(v.x >> 16), (int)(v.y & 0xffffu), (int)(v.y >> 16));
    int4 hi = make_int4((int)(v.z & 0xffffu), (int)(v.z >> 16), (int)(v.w & 0xffffu), (int)(v.w >> 16));
    reinterpret_cast<int4*>(dst)[2 * i] = lo;
    reinterpret_cast<int4*>(dst)[2 * i + 1] = hi;
  }
  if (blockIdx.x == 0)
    for (long long i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) dst[i] = (int)src[i];
}

inline int grid_for(const gcnb_ctx* ctx, long long work_items) {
  long long g = (work_items + kThreads - 1) / kThreads;
  const long long cap = (long long)ctx->sm_count * 16;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}
inline uint32_t thresh_of(float p) {
  if (p <= 0.f) return 0u;
  uint32_t t = dropout_threshold(p);
  return t == 0u ? 1u : t;
}
constexpr int kMaxReduceBlocks = 1184;  // 8 x 148

}  // namespace

int gcnb_highway_mix(gcnb_ctx* ctx, int n_rows, int hd, const float* H, int ldh, const float* T, int ldt,
                     const float* X, int ldx, float* Y, int ldy) {
  const int nf4 = (hd + 3) / 4;
  highway_mix_kernel<<<grid_for(ctx, (long long)n_rows * nf4), kThreads, 0, ctx->stream>>>(n_rows, nf4, H, ldh, T,
                                                                                            ldt, X, ldx, Y, ldy);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" int gcnb_highway_bwd_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t hd, int32_t ld, const float* dY,
                                    const float* X, const float* H, const float* T, int32_t act, float* dHpre,
                                    float* dTpre, float* dX) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, dY && X && H && T && dHpre && dTpre && dX, "null pointer");
  GCNB_REQUIRE(ctx, ld % 4 == 0 && ld >= ((hd + 3) / 4) * 4, "ld: multiple of 4, >= hd rounded to 4");
  if (n_rows == 0) return GCNB_OK;
  ProfScope scope(ctx, GCNB_TAG_ELEM);
  const int nf4 = (hd + 3) / 4;
  highway_bwd_kernel<<<grid_for(ctx, (long long)n_rows * nf4), kThreads, 0, ctx->stream>>>(n_rows, nf4, ld, dY, X, H,
                                                                                            T, act, dHpre, dTpre, dX);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" int gcnb_highway_bwd_bias_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t hd, int32_t ld, const float* dY,
                                         const float* X, const float* H, const float* T, int32_t act, float* dHpre,
                                         float* dTpre, float* dX, float* dbh, float* dbt) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, dY && X && H && T && dHpre && dTpre && dX && dbh && dbt, "null pointer");
  GCNB_REQUIRE(ctx, ld % 4 == 0 && ld >= ((hd + 3) / 4) * 4, "ld: multiple of 4, >= hd rounded to 4");
  const int nf4 = (hd + 3) / 4;
  if (nf4 > 128) {  // wider than one warp pass covers: the unfused kernels
    int rc = gcnb_highway_bwd_f32(ctx, n_rows, hd, ld, dY, X, H, T, act, dHpre, dTpre, dX);
    if (rc == GCNB_OK) rc = gcnb_colsum_f32(ctx, n_rows, hd, dHpre, ld, dbh, 0);
    if (rc == GCNB_OK) rc = gcnb_colsum_f32(ctx, n_rows, hd, dTpre, ld, dbt, 0);
    return rc;
  }
  const size_t need = 2 * gcnb_colsum_workspace_bytes(n_rows, hd);
  if (!ctx->ws || ctx->ws_bytes < need)
    return gcnb_fail(ctx, GCNB_E_WORKSPACE, "highway_bwd_bias needs %s%lld workspace bytes, have %lld", "",
                     (long long)need, (long long)ctx->ws_bytes);
  ProfScope scope(ctx, GCNB_TAG_ELEM);
  int blocks = cdiv(n_rows > 0 ? n_rows : 1, 64);
  if (blocks > kMaxReduceBlocks) blocks = kMaxReduceBlocks;
  float* partial = reinterpret_cast<float*>(ctx->ws);
  const int nch = (nf4 + 31) / 32;
  const PushPlan pp = gcnb_take_push(ctx, hd);
#define GCNB_LAUNCH_HWB(N)                                                                                         \
  highway_bwd_colsum_kernel<N><<<blocks, kThreads, 0, ctx->stream>>>(n_rows, nf4, ld, dY, X, H, T, act, dHpre, dTpre, \
                                                                     dX, partial, pp)
  if (nch <= 1) GCNB_LAUNCH_HWB(1);
  else if (nch == 2) GCNB_LAUNCH_HWB(2);
  else if (nch == 3) GCNB_LAUNCH_HWB(3);
  else GCNB_LAUNCH_HWB(4);
#undef GCNB_LAUNCH_HWB
  GCNB_LAUNCHED(ctx);
  colsum_stage2_kernel<<<cdiv(hd, 32), kThreads, 0, ctx->stream>>>(blocks, hd, (size_t)nf4 * 8, partial, dbh, 0);
  GCNB_LAUNCHED(ctx);
  colsum_stage2_kernel<<<cdiv(hd, 32), kThreads, 0, ctx->stream>>>(blocks, hd, (size_t)nf4 * 8, partial + (size_t)nf4 * 4, dbt, 0);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" int gcnb_act_bwd_bias_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t k, int32_t ld, const float* dY,
                                     const float* Yact, int32_t act, float dropout_p, uint64_t seed, int64_t row0,
                                     float* dZ, float* db) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, dY && Yact && dZ && db, "null pointer");
  GCNB_REQUIRE(ctx, ld % 4 == 0 && ld >= ((k + 3) / 4) * 4, "ld: multiple of 4, >= k rounded to 4");
  GCNB_REQUIRE(ctx, dropout_p >= 0.f && dropout_p < 1.f, "dropout_p in [0,1)");
  const int nf4 = (k + 3) / 4;
  if (nf4 > 128) {
    int rc = gcnb_act_bwd_f32(ctx, n_rows, k, ld, dY, Yact, act, dropout_p, seed, row0, dZ);
    if (rc == GCNB_OK) rc = gcnb_colsum_f32(ctx, n_rows, k, dZ, ld, db, 0);
    return rc;
  }
  const size_t need = gcnb_colsum_workspace_bytes(n_rows, k);
  if (!ctx->ws || ctx->ws_bytes < need)
    return gcnb_fail(ctx, GCNB_E_WORKSPACE, "act_bwd_bias needs %s%lld workspace bytes, have %lld", "",
                     (long long)need, (long long)ctx->ws_bytes);
  ProfScope scope(ctx, GCNB_TAG_ELEM);
  int blocks = cdiv(n_rows > 0 ? n_rows : 1, 64);
  if (blocks > kMaxReduceBlocks) blocks = kMaxReduceBlocks;
  float* partial = reinterpret_cast<float*>(ctx->ws);
  const float scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  const int nch = (nf4 + 31) / 32;
  const PushPlan pp = gcnb_take_push(ctx, k);
#define GCNB_LAUNCH_AB(N)                                                                                      \
  act_bwd_colsum_kernel<N><<<blocks, kThreads, 0, ctx->stream>>>(n_rows, nf4, ld, dY, Yact, act, thresh_of(dropout_p), \
                                                                 scale, seed, row0, dZ, partial, pp)
  if (nch <= 1) GCNB_LAUNCH_AB(1);
  else if (nch == 2) GCNB_LAUNCH_AB(2);
  else if (nch == 3) GCNB_LAUNCH_AB(3);
  else GCNB_LAUNCH_AB(4);
#undef GCNB_LAUNCH_AB
  GCNB_LAUNCHED(ctx);
  colsum_stage2_kernel<<<cdiv(k, 32), kThreads, 0, ctx->stream>>>(blocks, k, (size_t)nf4 * 4, partial, db, 0);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" int gcnb_act_bwd_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t k, int32_t ld, const float* dY,
                                const float* Yact, int32_t act, float dropout_p, uint64_t seed, int64_t row0,
                                float* dZ) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, dY && Yact && dZ, "null pointer");
  GCNB_REQUIRE(ctx, ld % 4 == 0 && ld >= ((k + 3) / 4) * 4, "ld: multiple of 4, >= k rounded to 4");
  GCNB_REQUIRE(ctx, dropout_p >= 0.f && dropout_p < 1.f, "dropout_p in [0,1)");
  if (n_rows == 0) return GCNB_OK;
  ProfScope scope(ctx, GCNB_TAG_ELEM);
  const int nf4 = (k + 3) / 4;
  const float scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  act_bwd_kernel<<<grid_for(ctx, (long long)n_rows * nf4), kThreads, 0, ctx->stream>>>(
      n_rows, nf4, ld, dY, Yact, act, thresh_of(dropout_p), scale, seed, row0, dZ);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" size_t gcnb_colsum_workspace_bytes(int32_t n_rows, int32_t k) {
  (void)n_rows;
  return (size_t)kMaxReduceBlocks * ((k + 3) / 4) * 4 * sizeof(float);
}

extern "C" int gcnb_colsum_f32(gcnb_ctx* ctx, int32_t n_rows, int32_t k, const float* A, int32_t lda, float* out,
                               int32_t accumulate) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, A && out, "null pointer");
  GCNB_REQUIRE(ctx, lda % 4 == 0 && lda >= ((k + 3) / 4) * 4, "lda: multiple of 4, >= k rounded to 4");
  const size_t need = gcnb_colsum_workspace_bytes(n_rows, k);
  if (!ctx->ws || ctx->ws_bytes < need)
    return gcnb_fail(ctx, GCNB_E_WORKSPACE, "colsum needs %s%lld workspace bytes, have %lld", "", (long long)need,
                     (long long)ctx->ws_bytes);
  ProfScope scope(ctx, GCNB_TAG_ELEM);
  const int nf4 = (k + 3) / 4;
  int blocks = cdiv(n_rows > 0 ? n_rows : 1, 64);
  if (blocks > kMaxReduceBlocks) blocks = kMaxReduceBlocks;
  float* partial = reinterpret_cast<float*>(ctx->ws);
  colsum_stage1_kernel<<<blocks, kThreads, 0, ctx->stream>>>(n_rows, nf4, A, lda, partial);
  GCNB_LAUNCHED(ctx);
  colsum_stage2_kernel<<<cdiv(k, 32), kThreads, 0, ctx->stream>>>(blocks, k, (size_t)nf4 * 4, partial, out, accumulate);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" int gcnb_xent_metrics_f32(gcnb_ctx* ctx, const float* P, int32_t ldp, int32_t n_classes,
                                     const int32_t* idx, const int32_t* labels, int32_t n_idx, float* metrics) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, P && metrics && (n_idx == 0 || (idx && labels)), "null pointer");
  if (n_idx == 0) return GCNB_OK;
  const size_t need = (size_t)kMaxReduceBlocks * 2 * sizeof(float);
  if (!ctx->ws || ctx->ws_bytes < need)
    return gcnb_fail(ctx, GCNB_E_WORKSPACE, "xent needs %s%lld workspace bytes, have %lld", "", (long long)need,
                     (long long)ctx->ws_bytes);
  ProfScope scope(ctx, GCNB_TAG_LOSS);
  int blocks = cdiv(n_idx, 8);
  if (blocks > kMaxReduceBlocks) blocks = kMaxReduceBlocks;
  float* partial = reinterpret_cast<float*>(ctx->ws);
  xent_metrics_stage1_kernel<<<blocks, kThreads, 0, ctx->stream>>>(P, ldp, n_classes, idx, labels, n_idx, partial);
  GCNB_LAUNCHED(ctx);
  xent_metrics_stage2_kernel<<<1, 32, 0, ctx->stream>>>(blocks, partial, metrics);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" int gcnb_xent_grad_f32(gcnb_ctx* ctx, const float* P, int32_t ldp, int32_t n_classes, int32_t n_rows,
                                  const int32_t* idx, const int32_t* labels, int32_t n_idx, float inv_n, float* G,
                                  int32_t ldg) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, P && G && (n_idx == 0 || (idx && labels)), "null pointer");
  ProfScope scope(ctx, GCNB_TAG_LOSS);
  GCNB_CUDA(ctx, cudaMemsetAsync(G, 0, (size_t)n_rows * ldg * sizeof(float), ctx->stream));
  if (n_idx == 0) return GCNB_OK;
  xent_grad_kernel<<<cdiv(n_idx, 8), kThreads, 0, ctx->stream>>>(P, ldp, n_classes, idx, labels, n_idx, inv_n, G, ldg);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" int gcnb_xent_grad_dense_f32(gcnb_ctx* ctx, const float* P, int32_t ldp, int32_t n_classes, int32_t n_rows,
                                        const int32_t* row_label, float inv_n, float* G, int32_t ldg) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, P && G && (n_rows == 0 || row_label), "null pointer");
  const int c4 = ((n_classes + 3) / 4) * 4;
  GCNB_REQUIRE(ctx, ldp % 4 == 0 && ldg % 4 == 0 && ldp >= c4 && ldg >= c4 && aligned16(P) && aligned16(G),
               "ldp / ldg: multiple of 4, >= classes rounded to 4; 16-byte aligned");
  if (n_rows == 0) return GCNB_OK;
  ProfScope scope(ctx, GCNB_TAG_LOSS);
  const PushPlan pp = gcnb_take_push(ctx, n_classes);
  int blocks = cdiv(n_rows, 8);
  const int cap = ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  xent_grad_dense_kernel<<<blocks, kThreads, 0, ctx->stream>>>(P, ldp, n_classes, n_rows, row_label, inv_n, G, ldg, pp);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" int gcnb_gather_argmax_f32(gcnb_ctx* ctx, const float* P, int32_t ldp, int32_t n_classes,
                                      const int32_t* idx, int32_t n_idx, int64_t* preds, float* probs) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, P && preds && (n_idx == 0 || idx), "null pointer");
  if (n_idx == 0) return GCNB_OK;
  ProfScope scope(ctx, GCNB_TAG_LOSS);
  gather_argmax_kernel<<<cdiv(n_idx, 8), kThreads, 0, ctx->stream>>>(P, ldp, n_classes, idx, n_idx,
                                                                      reinterpret_cast<long long*>(preds), probs);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" int gcnb_l1l2_f32(gcnb_ctx* ctx, const float* W, float* G, int64_t n, float coef, float* reg_sum) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, W && G && reg_sum, "null pointer");
  if (n == 0) return GCNB_OK;
  ProfScope scope(ctx, GCNB_TAG_ADAM);
  l1l2_kernel<<<grid_for(ctx, n), kThreads, 0, ctx->stream>>>(W, G, n, coef, reg_sum);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" int gcnb_adam_f32(gcnb_ctx* ctx, float* params, const float* grads, float* m, float* v, int64_t n,
                             float* state, float lr, float beta1, float beta2, float eps) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, params && grads && m && v && state, "null pointer");
  GCNB_REQUIRE(ctx, n % 4 == 0, "flat parameter buffer length must be a multiple of 4");
  GCNB_REQUIRE(ctx, aligned16(params) && aligned16(grads) && aligned16(m) && aligned16(v), "16-byte alignment");
  ProfScope scope(ctx, GCNB_TAG_ADAM);
  adam_tick_kernel<<<1, 1, 0, ctx->stream>>>(state, lr, beta1, beta2);
  GCNB_LAUNCHED(ctx);
  if (n > 0) {
    adam_kernel<<<grid_for(ctx, n / 4), kThreads, 0, ctx->stream>>>(
        reinterpret_cast<float4*>(params), reinterpret_cast<const float4*>(grads), reinterpret_cast<float4*>(m),
        reinterpret_cast<float4*>(v), n / 4, state, beta1, beta2, eps);
    GCNB_LAUNCHED(ctx);
  }
  return GCNB_OK;
}

extern "C" int gcnb_dropout_mask_u8(gcnb_ctx* ctx, int32_t n_rows, int32_t k, float p, uint64_t seed, int64_t row0,
                                    uint8_t* mask) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, mask, "null pointer");
  GCNB_REQUIRE(ctx, p >= 0.f && p < 1.f, "p in [0,1)");
  if (n_rows == 0 || k == 0) return GCNB_OK;
  ProfScope scope(ctx, GCNB_TAG_ELEM);
  const long long total = (long long)n_rows * ((k + 3) / 4);
  dropout_mask_kernel<<<grid_for(ctx, total), kThreads, 0, ctx->stream>>>(n_rows, k, thresh_of(p), seed, row0, mask);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

static int move_rows(gcnb_ctx* ctx, const float* src, int ld_src, const int* idx, int n_idx, int k, float* dst,
                     int ld_dst, int scatter) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, src && dst && (n_idx == 0 || idx), "null pointer");
  const int k4 = ((k + 3) / 4) * 4;
  GCNB_REQUIRE(ctx, ld_src % 4 == 0 && ld_dst % 4 == 0 && ld_src >= k4 && ld_dst >= k4, "ld: multiple of 4, >= k rounded to 4");
  if (n_idx == 0 || k == 0) return GCNB_OK;
  ProfScope scope(ctx, GCNB_TAG_ELEM);
  const int nf4 = k4 / 4;
  move_rows_kernel<<<grid_for(ctx, (long long)n_idx * nf4), kThreads, 0, ctx->stream>>>(src, ld_src, idx, n_idx, nf4,
                                                                                        dst, ld_dst, scatter);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}
extern "C" int gcnb_csr_to_dense_f32(gcnb_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const float* val,
                                     int32_t n_rows, int32_t n_cols, float* dense, int32_t ld) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, n_rows >= 0 && n_cols >= 0 && ld >= n_cols, "bad shape");
  if (n_rows == 0 || n_cols == 0) return GCNB_OK;
  GCNB_REQUIRE(ctx, rowptr && dense, "null pointer");
  ProfScope scope(ctx, GCNB_TAG_COPY);
  GCNB_CUDA(ctx, cudaMemsetAsync(dense, 0, (size_t)n_rows * ld * sizeof(float), ctx->stream));
  csr_to_dense_kernel<<<cdiv(n_rows, 8), kThreads, 0, ctx->stream>>>(rowptr, colidx, val, n_rows, dense, ld);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}

extern "C" int gcnb_gather_rows_f32(gcnb_ctx* ctx, const float* src, int32_t ld_src, const int32_t* idx, int32_t n_idx,
                                    int32_t k, float* dst, int32_t ld_dst) {
  return move_rows(ctx, src, ld_src, idx, n_idx, k, dst, ld_dst, 0);
}
extern "C" int gcnb_scatter_rows_f32(gcnb_ctx* ctx, const float* src, int32_t ld_src, const int32_t* idx,
                                     int32_t n_idx, int32_t k, float* dst, int32_t ld_dst) {
  return move_rows(ctx, src, ld_src, idx, n_idx, k, dst, ld_dst, 1);
}

extern "C" int gcnb_expand_u16_i32(gcnb_ctx* ctx, const uint16_t* src, int64_t n, int32_t* dst) {
  if (!ctx) return GCNB_E_INVALID;
  if (n == 0) return GCNB_OK;
  GCNB_REQUIRE(ctx, src && dst && n > 0, "null pointer");
  GCNB_REQUIRE(ctx, aligned16(src) && aligned16(dst), "16-byte alignment");
  ProfScope scope(ctx, GCNB_TAG_COPY);
  expand_u16_kernel<<<grid_for(ctx, (n + 7) / 8), kThreads, 0, ctx->stream>>>(src, n, dst);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}
