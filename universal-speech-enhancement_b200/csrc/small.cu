// Small fp32 kernels: time-embedding MLP, the per-ResBlock Dense_0 biases, and the bottleneck attention
// block.  Reference: ncsnpp.py:349-368 (temb), layerspp.py:302-303 (Dense_0(act(temb))),
// layerspp.py:77-93 + layers.py:639-650 (AttnBlockpp / NIN).  Together < 0.01 % of the FLOPs.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace use {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one block per sample; warp-per-output-row GEMVs with coalesced weight reads
__global__ void __launch_bounds__(512) temb_mlp_kernel(const float* __restrict__ gfp, int gfp_bstride,
                                                        const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, float* __restrict__ out, int nf) {
  extern __shared__ float sm[];  // in[2nf], h1[4nf]
  float* sin_ = sm;
  float* h1 = sm + 2 * nf;
  const int b = blockIdx.x, K1 = 2 * nf, D = 4 * nf;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < K1; i += blockDim.x) sin_[i] = gfp[static_cast<size_t>(b) * gfp_bstride + i];
  __syncthreads();
  for (int n = warp; n < D; n += nw) {
    float acc = 0.f;
    for (int k = lane; k < K1; k += 32) acc += w1[n * K1 + k] * sin_[k];
    acc = warp_sum(acc);
    if (lane == 0) h1[n] = silu(acc + b1[n]);  // act(temb) between the two Linear layers
  }
  __syncthreads();
  for (int n = warp; n < D; n += nw) {
    float acc = 0.f;
    for (int k = lane; k < D; k += 32) acc += w2[n * D + k] * h1[k];
    acc = warp_sum(acc);
    if (lane == 0) out[b * D + n] = silu(acc + b2[n]);  // every consumer applies act(temb) first
  }
}

void launch_temb_mlp(const float* gfp, int gfp_bstride, const float* w1, const float* b1, const float* w2, const float* b2,
                     float* out, int B, int nf, cudaStream_t st) {
  temb_mlp_kernel<<<B, 512, 6 * nf * sizeof(float), st>>>(gfp, gfp_bstride, w1, b1, w2, b2, out, nf);
}

// out[b][n] = base[n] + W[n][:] . temb[b][:]   (all Dense_0 of all ResBlocks stacked along n)
__global__ void __launch_bounds__(256) dense_all_kernel(const float* __restrict__ temb, const float* __restrict__ W,
                                                         const float* __restrict__ base, float* __restrict__ out, int B,
                                                         int rows, int K) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* wr = W + static_cast<size_t>(warp) * K;
  const float bs = base[warp];
  for (int b = 0; b < B; ++b) {
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc += wr[k] * temb[b * K + k];
    acc = warp_sum(acc);
    if (lane == 0) out[static_cast<size_t>(b) * rows + warp] = acc + bs;
  }
}

void launch_dense_all(const float* temb, const float* W, const float* base, float* out, int B, int rows, int K,
                      cudaStream_t st) {
  const int blocks = (rows * 32 + 255) / 256;
  dense_all_kernel<<<blocks, 256, 0, st>>>(temb, W, base, out, B, rows, K);
}

// out[m][n] = in[m][:] . W[:][n] + b[n]   (NIN: W is [in][out])
constexpr int LIN_RM = 8;
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                      const float* __restrict__ b, float* __restrict__ out, int M, int K,
                                                      int N) {
  extern __shared__ float sin_[];  // [LIN_RM][K]
  const int m0 = blockIdx.x * LIN_RM;
  for (int i = threadIdx.x; i < LIN_RM * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    sin_[i] = (m0 + r < M) ? in[static_cast<size_t>(m0 + r) * K + k] : 0.f;
  }
  __syncthreads();
  for (int n = blockIdx.y * blockDim.x + threadIdx.x; n < N; n += gridDim.y * blockDim.x) {
    float acc[LIN_RM];
#pragma unroll
    for (int r = 0; r < LIN_RM; ++r) acc[r] = 0.f;
    for (int k = 0; k < K; ++k) {
      const float w = W[static_cast<size_t>(k) * N + n];
#pragma unroll
      for (int r = 0; r < LIN_RM; ++r) acc[r] += sin_[r * K + k] * w;
    }
    const float bb = b[n];
#pragma unroll
    for (int r = 0; r < LIN_RM; ++r)
      if (m0 + r < M) out[static_cast<size_t>(m0 + r) * N + n] = acc[r] + bb;
  }
}

void launch_linear(const float* in, const float* W, const float* b, float* out, int M, int K, int N, cudaStream_t st) {
  dim3 grid((M + LIN_RM - 1) / LIN_RM, (N + 255) / 256);
  linear_kernel<<<grid, 256, LIN_RM * K * sizeof(float), st>>>(in, W, b, out, M, K, N);
}

// softmax(q k^T / sqrt(C)) v over all P = H*W positions, one block per (sample, query position)
__global__ void __launch_bounds__(256) attn_core_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                         const float* __restrict__ v, float* __restrict__ out, int P,
                                                         int C) {
  extern __shared__ float sm[];  // q[C], w[P], red[32]
  float* sq = sm;
  float* sw = sm + C;
  float* red = sw + P;
  const int b = blockIdx.y, i = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* qb = q + (static_cast<size_t>(b) * P + i) * C;
  const float* kb = k + static_cast<size_t>(b) * P * C;
  const float* vb = v + static_cast<size_t>(b) * P * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sq[c] = qb[c];
  __syncthreads();
  const float scl = rsqrtf(static_cast<float>(C));
  for (int j = warp; j < P; j += nw) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc += sq[c] * kb[static_cast<size_t>(j) * C + c];
    acc = warp_sum(acc);
    if (lane == 0) sw[j] = acc * scl;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < P; j += blockDim.x) mx = fmaxf(mx, sw[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < nw; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < P; j += blockDim.x) {
    const float e = expf(sw[j] - mx);
    sw[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < nw; ++w) sum += red[w];
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < P; ++j) acc += sw[j] * vb[static_cast<size_t>(j) * C + c];
    out[(static_cast<size_t>(b) * P + i) * C + c] = acc * inv;
  }
}

void launch_attn_core(const float* q, const float* k, const float* v, float* out, int B, int P, int C, cudaStream_t st) {
  dim3 grid(P, B);
  attn_core_kernel<<<grid, 256, (C + P + 32) * sizeof(float), st>>>(q, k, v, out, P, C);
}

template <typename T>
__global__ void __launch_bounds__(256) add_scale_kernel(const T* __restrict__ x, const float* __restrict__ h, float scale,
                                                         T* __restrict__ out, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = static_cast<T>((static_cast<float>(x[i]) + h[i]) * scale);
}
template <typename T>
__global__ void __launch_bounds__(256) act_to_f32_kernel(const T* __restrict__ x, float* __restrict__ out, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = static_cast<float>(x[i]);
}

void launch_add_scale(int dt, const void* x, const float* h, float scale, void* out, size_t n, cudaStream_t st) {
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
  if (dt == kBF16)
    add_scale_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, h, scale, (__nv_bfloat16*)out, n);
  else
    add_scale_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, h, scale, (float*)out, n);
}
void launch_act_to_f32(int dt, const void* x, float* out, size_t n, cudaStream_t st) {
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
  if (dt == kBF16)
    act_to_f32_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, out, n);
  else
    act_to_f32_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, out, n);
}

}  // namespace use
