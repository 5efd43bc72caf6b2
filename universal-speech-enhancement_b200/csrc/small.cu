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

// out[b][n] = base[n] + W[n][:] . temb[b][:]   (all Dense_0 of all ResBlocks stacked along n).  temb_bstride = 0: every
// sample shares one time (the sampling loop): the row is computed once and written B times.
__global__ void __launch_bounds__(256) dense_all_kernel(const float* __restrict__ temb, int temb_bstride,
                                                         const float* __restrict__ W, const float* __restrict__ base,
                                                         float* __restrict__ out, int B, int rows, int K) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* wr = W + static_cast<size_t>(warp) * K;
  const float bs = base[warp];
  float last = 0.f;
  for (int b = 0; b < B; ++b) {
    if (b == 0 || temb_bstride != 0) {
      float acc = 0.f;
      for (int k = lane; k < K; k += 32) acc += wr[k] * temb[static_cast<size_t>(b) * temb_bstride + k];
      last = warp_sum(acc) + bs;
    }
    if (lane == 0) out[static_cast<size_t>(b) * rows + warp] = last;
  }
}

void launch_dense_all(const float* temb, int temb_bstride, const float* W, const float* base, float* out, int B, int rows,
                      int K, cudaStream_t st) {
  const int blocks = (rows * 32 + 255) / 256;
  dense_all_kernel<<<blocks, 256, 0, st>>>(temb, temb_bstride, W, base, out, B, rows, K);
}

// NIN layers of the attention block: out[m][n] = in[m][:] . W[:][n] + b[n]   (W is [in][out], layers.py:639-650).
// A block owns NIN_RM rows x NIN_CN columns; its 256 threads are (column, K quarter): four partial dot products per
// output run in parallel and are summed in a fixed order (deterministic).  At batch 1 the 80 x 256 x 256 product is
// spread over 10 x 4 (x 3 for q, k, v) blocks with 64-step loops; the round-1 kernel walked K = 256 serially in 10 blocks
// and the seven launches of the attention block cost 0.2 ms per evaluation (3.5 % of a batch-1 step).
constexpr int NIN_RM = 8;
constexpr int NIN_CN = 64;
constexpr int NIN_KP = 4;

struct NinJob {
  const float* W;
  const float* b;
  float* out;
};

template <typename TIn>
__device__ __forceinline__ void nin_block_partial(const TIn* __restrict__ in, const float* __restrict__ W, int M, int K, int N,
                                                  int m0, int n, int kp, float* sin_, float (&acc)[NIN_RM]) {
  for (int i = threadIdx.x; i < NIN_RM * K; i += blockDim.x) {
    const int r = i / K, k = i - r * K;
    sin_[i] = (m0 + r < M) ? static_cast<float>(in[static_cast<size_t>(m0 + r) * K + k]) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < NIN_RM; ++r) acc[r] = 0.f;
  const int kq = (K + NIN_KP - 1) / NIN_KP;
  const int k0 = kp * kq, k1 = min(K, k0 + kq);
  if (n < N) {
#pragma unroll 8
    for (int k = k0; k < k1; ++k) {
      const float w = __ldg(W + static_cast<size_t>(k) * N + n);
#pragma unroll
      for (int r = 0; r < NIN_RM; ++r) acc[r] = fmaf(sin_[r * K + k], w, acc[r]);
    }
  }
}

// q, k, v = NIN_0/1/2(GroupNorm(x)): blockIdx.z selects the layer; the input is read in the activation dtype
template <typename TIn>
__global__ void __launch_bounds__(NIN_CN * NIN_KP) nin_qkv_kernel(const TIn* __restrict__ in, NinJob j0, NinJob j1, NinJob j2,
                                                                   int M, int K, int N) {
  extern __shared__ float nsm[];  // in[NIN_RM][K], part[NIN_KP][NIN_RM][NIN_CN]
  float* sin_ = nsm;
  float* part = nsm + NIN_RM * K;
  const NinJob J = blockIdx.z == 0 ? j0 : (blockIdx.z == 1 ? j1 : j2);
  const int m0 = blockIdx.x * NIN_RM;
  const int nl = threadIdx.x % NIN_CN, kp = threadIdx.x / NIN_CN;
  const int n = blockIdx.y * NIN_CN + nl;
  float acc[NIN_RM];
  nin_block_partial<TIn>(in, J.W, M, K, N, m0, n, kp, sin_, acc);
#pragma unroll
  for (int r = 0; r < NIN_RM; ++r) part[(kp * NIN_RM + r) * NIN_CN + nl] = acc[r];
  __syncthreads();
  for (int i = threadIdx.x; i < NIN_RM * NIN_CN; i += blockDim.x) {
    const int r = i / NIN_CN, c = i - r * NIN_CN;
    const int nn = blockIdx.y * NIN_CN + c;
    if (m0 + r < M && nn < N) {
      float v = part[r * NIN_CN + c];
#pragma unroll
      for (int q = 1; q < NIN_KP; ++q) v += part[(q * NIN_RM + r) * NIN_CN + c];
      J.out[static_cast<size_t>(m0 + r) * N + nn] = v + __ldg(J.b + nn);
    }
  }
}

// out = (x + NIN_3(att)) * scale in the activation dtype (layerspp.py:89-93)
template <typename T>
__global__ void __launch_bounds__(NIN_CN * NIN_KP) nin_proj_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                                    const float* __restrict__ b, const T* __restrict__ x,
                                                                    float scale, T* __restrict__ out, int M, int K, int N) {
  extern __shared__ float nsm[];
  float* sin_ = nsm;
  float* part = nsm + NIN_RM * K;
  const int m0 = blockIdx.x * NIN_RM;
  const int nl = threadIdx.x % NIN_CN, kp = threadIdx.x / NIN_CN;
  const int n = blockIdx.y * NIN_CN + nl;
  float acc[NIN_RM];
  nin_block_partial<float>(in, W, M, K, N, m0, n, kp, sin_, acc);
#pragma unroll
  for (int r = 0; r < NIN_RM; ++r) part[(kp * NIN_RM + r) * NIN_CN + nl] = acc[r];
  __syncthreads();
  for (int i = threadIdx.x; i < NIN_RM * NIN_CN; i += blockDim.x) {
    const int r = i / NIN_CN, c = i - r * NIN_CN;
    const int nn = blockIdx.y * NIN_CN + c;
    if (m0 + r < M && nn < N) {
      float v = part[r * NIN_CN + c];
#pragma unroll
      for (int q = 1; q < NIN_KP; ++q) v += part[(q * NIN_RM + r) * NIN_CN + c];
      const size_t o = static_cast<size_t>(m0 + r) * N + nn;
      out[o] = static_cast<T>((static_cast<float>(x[o]) + (v + __ldg(b + nn))) * scale);
    }
  }
}

static size_t nin_smem(int K) { return (static_cast<size_t>(NIN_RM) * K + NIN_KP * NIN_RM * NIN_CN) * sizeof(float); }

void launch_nin_qkv(int dt, const void* in, const float* W0, const float* b0, float* q, const float* W1, const float* b1,
                    float* k, const float* W2, const float* b2, float* v, int M, int C, cudaStream_t st) {
  dim3 grid((M + NIN_RM - 1) / NIN_RM, (C + NIN_CN - 1) / NIN_CN, 3);
  const NinJob j0{W0, b0, q}, j1{W1, b1, k}, j2{W2, b2, v};
  if (dt == kBF16)
    nin_qkv_kernel<__nv_bfloat16><<<grid, NIN_CN * NIN_KP, nin_smem(C), st>>>((const __nv_bfloat16*)in, j0, j1, j2, M, C, C);
  else
    nin_qkv_kernel<float><<<grid, NIN_CN * NIN_KP, nin_smem(C), st>>>((const float*)in, j0, j1, j2, M, C, C);
}

void launch_nin_proj(int dt, const float* att, const float* W, const float* b, const void* x, float scale, void* out, int M,
                     int C, cudaStream_t st) {
  dim3 grid((M + NIN_RM - 1) / NIN_RM, (C + NIN_CN - 1) / NIN_CN);
  if (dt == kBF16)
    nin_proj_kernel<__nv_bfloat16><<<grid, NIN_CN * NIN_KP, nin_smem(C), st>>>(att, W, b, (const __nv_bfloat16*)x, scale,
                                                                              (__nv_bfloat16*)out, M, C, C);
  else
    nin_proj_kernel<float><<<grid, NIN_CN * NIN_KP, nin_smem(C), st>>>(att, W, b, (const float*)x, scale, (float*)out, M, C, C);
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
#pragma unroll 8
    for (int j = 0; j < P; ++j) acc += sw[j] * vb[static_cast<size_t>(j) * C + c];  // (same order; the loads of 8 rows in flight)
    out[(static_cast<size_t>(b) * P + i) * C + c] = acc * inv;
  }
}

void launch_attn_core(const float* q, const float* k, const float* v, float* out, int B, int P, int C, cudaStream_t st) {
  dim3 grid(P, B);
  attn_core_kernel<<<grid, 256, (C + P + 32) * sizeof(float), st>>>(q, k, v, out, P, C);
}

}  // namespace use
