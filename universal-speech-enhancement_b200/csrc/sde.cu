// SDE arithmetic of the sampler, fused with the network's output head.
// Reference: NCSNpp.forward tail (ncsnpp.py:483-500: h / t, output_layer 1x1 4->2, view_as_complex),
// ScoreModel.forward_score sign (model_wrapper.py:137), RSDE.discretize + OUVESDE.sde (sdes.py:75-92,
// 159-173,216-224), ReverseDiffusionPredictor.update_fn (predictors.py:61-68), OUVESDE.prior_sampling
// (sdes.py:248-254).  In-kernel noise: Philox4x32-10 + Box-Muller, one stream per (seed, step, global clip
// index, element) so results do not depend on how clips are sharded over GPUs.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace use {

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// complex standard normal (Re, Im each variance 1/2): what torch.randn_like(complex64) draws
__device__ __forceinline__ float2 philox_cnormal(unsigned long long seed, uint32_t step, uint32_t clip, uint64_t elem) {
  uint32_t c[4] = {static_cast<uint32_t>(elem), static_cast<uint32_t>(elem >> 32), step, clip};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const float u1 = (static_cast<float>(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  const float u2 = (static_cast<float>(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r = sqrtf(-logf(u1));  // sqrt(-2 ln u1) / sqrt(2)
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}

// xr: fp32 input pyramid level 0.  PC = 4: [n][4] = [Re x, Im x, Re Y, Im Y]; PC = 2: [n][2] = [Re x, Im x]
// (discriminative network, no conditioning); PC = 6 (condition="both", model_wrapper.py:43-46,287-288): the 4-channel
// block [n][4] followed by a 2-channel block [n][2] = [Re Y2, Im Y2] (the engine keeps 6-channel pyramids as a 4 + 2
// pair so that every 4- / 2-channel kernel is reused).  xpad (optional): act dtype [n][128 B of channels], channels
// 0..PC-1 = the same values as MMA operands, the rest zero -- the tcgen05 input convolution reads it as one K chunk.
template <typename T, int PC>
__global__ void __launch_bounds__(256) pack_input_kernel(const float2* __restrict__ x, const float2* __restrict__ Y,
                                                          const float2* __restrict__ Y2, float* __restrict__ xr,
                                                          T* __restrict__ xpad, size_t n) {
  constexpr int V = Vec<T>::N;
  constexpr int VPP = 8;  // 16-byte vectors per padded pixel (128 B)
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n * VPP;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = i / VPP;
    const int v = static_cast<int>(i % VPP);
    float f[V];
#pragma unroll
    for (int j = 0; j < V; ++j) f[j] = 0.f;
    if (v == 0) {
      const float2 a = x[pix];
      f[0] = 2.f * a.x - 1.0f;
      f[1] = 2.f * a.y - 1.0f;
      if constexpr (PC >= 4) {
        const float2 b = Y[pix];
        f[2] = 2.f * b.x - 1.0f;
        f[3] = 2.f * b.y - 1.0f;
        reinterpret_cast<float4*>(xr)[pix] = make_float4(f[0], f[1], f[2], f[3]);
      } else {
        reinterpret_cast<float2*>(xr)[pix] = make_float2(f[0], f[1]);
      }
    }
    if constexpr (PC == 6) {
      // channels 4, 5: vector 0 of a bf16 pixel (8 channels per vector), vector 1 of an fp32 pixel (4 per vector)
      if (v == (V == 8 ? 0 : 1)) {
        const float2 c = Y2[pix];
        const float g0 = 2.f * c.x - 1.0f, g1 = 2.f * c.y - 1.0f;
        f[4 % V] = g0;
        f[5 % V] = g1;
        reinterpret_cast<float2*>(xr + n * 4)[pix] = make_float2(g0, g1);
      }
    }
    if (xpad != nullptr) Vec<T>::store_operand(xpad + pix * (VPP * V) + v * V, f);
  }
}

void launch_pack_input(int dt, int pc, const float2* x, const float2* Y, const float2* Y2, float* xr, void* xpad, size_t n,
                       cudaStream_t st) {
  const int blocks = static_cast<int>(std::min<size_t>((n * 8 + 255) / 256, 148 * 32));
  if (dt == kBF16) {
    if (pc == 6) pack_input_kernel<__nv_bfloat16, 6><<<blocks, 256, 0, st>>>(x, Y, Y2, xr, (__nv_bfloat16*)xpad, n);
    else if (pc == 4) pack_input_kernel<__nv_bfloat16, 4><<<blocks, 256, 0, st>>>(x, Y, Y2, xr, (__nv_bfloat16*)xpad, n);
    else pack_input_kernel<__nv_bfloat16, 2><<<blocks, 256, 0, st>>>(x, Y, Y2, xr, (__nv_bfloat16*)xpad, n);
  } else {
    if (pc == 6) pack_input_kernel<float, 6><<<blocks, 256, 0, st>>>(x, Y, Y2, xr, (float*)xpad, n);
    else if (pc == 4) pack_input_kernel<float, 4><<<blocks, 256, 0, st>>>(x, Y, Y2, xr, (float*)xpad, n);
    else pack_input_kernel<float, 2><<<blocks, 256, 0, st>>>(x, Y, Y2, xr, (float*)xpad, n);
  }
}

__global__ void __launch_bounds__(256) final_step_kernel(StepArgs a) {
  const size_t n = a.per_clip * a.B;
  float w0[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, w1[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < a.pc; ++j) { w0[j] = a.ow[j]; w1[j] = a.ow[a.pc + j]; }
  const float b0 = a.ob[0], b1 = a.ob[1];
  const float G2 = a.G * a.G;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / a.per_clip);
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 p2 = make_float2(0.f, 0.f);  // channels 4, 5 of a 6-channel pyramid (kept as a separate 2-channel tensor)
    if (a.pc >= 4) {
      p = __ldg(reinterpret_cast<const float4*>(a.pyramid) + i);
      if (a.pc == 6) p2 = __ldg(reinterpret_cast<const float2*>(a.pyramid2) + i);
    } else {
      const float2 q = __ldg(reinterpret_cast<const float2*>(a.pyramid) + i);
      p.x = q.x; p.y = q.y;
    }
    if (a.t != nullptr) {  // scale_by_sigma: divide by the TIME value (ncsnpp.py:492-494)
      const float t = a.t[static_cast<size_t>(b) * a.t_bstride];
      p.x /= t; p.y /= t; p.z /= t; p.w /= t;
      p2.x /= t; p2.y /= t;
    }
    float ore = b0 + w0[0] * p.x + w0[1] * p.y + w0[2] * p.z + w0[3] * p.w;
    float oim = b1 + w1[0] * p.x + w1[1] * p.y + w1[2] * p.z + w1[3] * p.w;
    if (a.pc == 6) {
      ore += w0[4] * p2.x + w0[5] * p2.y;
      oim += w1[4] * p2.x + w1[5] * p2.y;
    }
    const float2 score = make_float2(a.out_sign * ore, a.out_sign * oim);
    if (a.score != nullptr) a.score[i] = score;
    if (a.x != nullptr) {
      const float2 x = a.x[i], Y = a.Y[i];
      float2 z;
      if (a.z != nullptr) z = a.z[i];
      else z = philox_cnormal(a.seed, a.step, a.clip0 + b, i - static_cast<size_t>(b) * a.per_clip);
      float2 xm;
      if (a.mode == kStepDrift) {
        a.x_mean[i] = make_float2(a.theta * (Y.x - x.x) + (-G2 * score.x) * a.pf, a.theta * (Y.y - x.y) + (-G2 * score.y) * a.pf);
        continue;
      }
      if (a.mode == kStepEulerMaruyama) {
        // f = theta (Y - x) - g^2 score pf ; x_mean = x + f (-1/N) ; x = x_mean + g sqrt(1/N) z   (predictors.py:40-53,
        // RSDE.rsde_parts sdes.py:128-150); a.G = g(t_i) here, a.Gz = g sqrt(1/N) (0 for the probability flow)
        const float dre = a.theta * (Y.x - x.x) + (-G2 * score.x) * a.pf, dim_ = a.theta * (Y.y - x.y) + (-G2 * score.y) * a.pf;
        xm = make_float2(x.x + dre * (-a.dt), x.y + dim_ * (-a.dt));
      } else {
        // f = theta (Y - x) dt ; rev_f = f - G^2 score pf ; x_mean = x - rev_f ; x = x_mean + G z
        const float fre = (a.theta * (Y.x - x.x)) * a.dt, fim = (a.theta * (Y.y - x.y)) * a.dt;
        const float rre = fre - G2 * score.x * a.pf, rim = fim - G2 * score.y * a.pf;
        xm = make_float2(x.x - rre, x.y - rim);
      }
      a.x_mean[i] = xm;
      a.x_next[i] = make_float2(xm.x + a.Gz * z.x, xm.y + a.Gz * z.y);
    }
  }
}

// ---- correctors (sampling/correctors.py:37-98) ---------------------------------------------------------------
// Langevin: the step size couples the whole batch through two batch-mean norms (correctors.py:55-57).  Deterministic
// two-pass reduction: pass 1 = per-(clip, block) sums of |grad|^2 and |noise|^2 in a fixed order (double), pass 2 = one
// block folds the partials in index order, takes sqrt per clip, the batch means, and writes the step coefficients.
constexpr int kRedBlocks = 64;  // partial sums per clip

__global__ void __launch_bounds__(256) sumsq_pair_kernel(const float2* __restrict__ grad, const float2* __restrict__ z,
                                                          double* __restrict__ part, unsigned long long seed,
                                                          unsigned int draw, unsigned int clip0, size_t per_clip) {
  const int b = blockIdx.y, blk = blockIdx.x;
  const size_t chunk = (per_clip + kRedBlocks - 1) / kRedBlocks;
  const size_t lo = blk * chunk, hi = lo + chunk < per_clip ? lo + chunk : per_clip;
  double sg = 0.0, sz = 0.0;
  for (size_t i = lo + threadIdx.x; i < hi; i += 256) {
    const float2 g = grad[static_cast<size_t>(b) * per_clip + i];
    const float2 n = z ? z[static_cast<size_t>(b) * per_clip + i] : philox_cnormal(seed, draw, clip0 + b, i);
    sg += static_cast<double>(g.x) * g.x + static_cast<double>(g.y) * g.y;
    sz += static_cast<double>(n.x) * n.x + static_cast<double>(n.y) * n.y;
  }
  __shared__ double sh[2][256];
  sh[0][threadIdx.x] = sg;
  sh[1][threadIdx.x] = sz;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {  // fixed tree: deterministic
    if (threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[(static_cast<size_t>(b) * kRedBlocks + blk) * 2] = sh[0][0];
    part[(static_cast<size_t>(b) * kRedBlocks + blk) * 2 + 1] = sh[1][0];
  }
}

// coef[0] = step_size = 2 (snr * mean_b ||z_b|| / mean_b ||grad_b||)^2 ; coef[1] = sqrt(2 step_size)
__global__ void langevin_coef_kernel(const double* __restrict__ part, float* __restrict__ coef, float snr, int B) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float gn = 0.f, zn = 0.f;
  for (int b = 0; b < B; ++b) {
    double sg = 0.0, sz = 0.0;
    for (int k = 0; k < kRedBlocks; ++k) {
      sg += part[(static_cast<size_t>(b) * kRedBlocks + k) * 2];
      sz += part[(static_cast<size_t>(b) * kRedBlocks + k) * 2 + 1];
    }
    gn += static_cast<float>(sqrt(sg));
    zn += static_cast<float>(sqrt(sz));
  }
  gn /= static_cast<float>(B);
  zn /= static_cast<float>(B);
  const float r = snr * zn / gn;
  const float step = r * r * 2.f;
  coef[0] = step;
  coef[1] = sqrtf(step * 2.f);
}

// x_mean = x + step * grad ; x = x_mean + noise * sqrt(2 step)   (correctors.py:61-62,95-96); in place on x allowed
__global__ void __launch_bounds__(256) corrector_update_kernel(const float2* __restrict__ x, const float2* __restrict__ grad,
                                                                const float2* __restrict__ z, const float* __restrict__ coef_dev,
                                                                float step_imm, float nz_imm, float2* __restrict__ x_mean,
                                                                float2* __restrict__ x_next, unsigned long long seed,
                                                                unsigned int draw, unsigned int clip0, size_t per_clip, size_t n) {
  const float step = coef_dev ? coef_dev[0] : step_imm;
  const float nz = coef_dev ? coef_dev[1] : nz_imm;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t b = static_cast<uint32_t>(i / per_clip);
    const float2 zz = z ? z[i] : philox_cnormal(seed, draw, clip0 + b, i - static_cast<size_t>(b) * per_clip);
    const float2 xv = x[i], g = grad[i];
    const float2 xm = make_float2(xv.x + step * g.x, xv.y + step * g.y);
    x_mean[i] = xm;
    x_next[i] = make_float2(xm.x + zz.x * nz, xm.y + zz.y * nz);
  }
}

size_t corrector_scratch_bytes(int B) { return static_cast<size_t>(B) * kRedBlocks * 2 * sizeof(double) + 64; }

void launch_corrector_step(const CorrectorArgs& c, cudaStream_t st) {
  const size_t n = c.per_clip * c.B;
  const float* coef_dev = nullptr;
  if (c.langevin) {
    double* part = reinterpret_cast<double*>(c.scratch);
    float* coef = reinterpret_cast<float*>(reinterpret_cast<char*>(c.scratch) + static_cast<size_t>(c.B) * kRedBlocks * 2 * sizeof(double));
    sumsq_pair_kernel<<<dim3(kRedBlocks, c.B), 256, 0, st>>>(c.grad, c.z, part, c.seed, c.draw, c.clip0, c.per_clip);
    langevin_coef_kernel<<<1, 32, 0, st>>>(part, coef, c.snr, c.B);
    coef_dev = coef;
  }
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
  corrector_update_kernel<<<blocks, 256, 0, st>>>(c.x, c.grad, c.z, coef_dev, c.step, sqrtf(c.step * 2.f), c.x_mean, c.x_next,
                                                   c.seed, c.draw, c.clip0, c.per_clip, n);
}

// ---- forward half of ScoreModel.train_step (model_wrapper.py:147-208) ----------------------------------------------
// perturbed = exp(-theta t) x0 + (1 - exp(-theta t)) y + std(t) z   (OUVESDE.marginal_prob, sdes.py:225-247); the
// per-sample coefficients e_b = exp(-theta t_b) and std_b come from the host (the reference's float32 torch expressions)
__global__ void __launch_bounds__(256) perturb_kernel(const float2* __restrict__ X0, const float2* __restrict__ Y,
                                                       const float2* __restrict__ z, const float* __restrict__ coef,
                                                       float2* __restrict__ xt, unsigned long long seed, unsigned int clip0,
                                                       size_t per_clip, size_t n, int B) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t b = static_cast<uint32_t>(i / per_clip);
    const float e = coef[b], sd = coef[B + b];
    const float2 zz = z ? z[i] : philox_cnormal(seed, 0xfffffffeu, clip0 + b, i - static_cast<size_t>(b) * per_clip);
    const float2 x0 = X0[i], y = Y[i];
    const float om = 1.f - e;
    xt[i] = make_float2((e * x0.x + om * y.x) + sd * zz.x, (e * x0.y + om * y.y) + sd * zz.y);
  }
}

// err = score std + z ; per-(clip, block) partial sums of |err|^2 (mse) or |err| (mae), fixed order (model_wrapper.py:124-133)
__global__ void __launch_bounds__(256) dsm_partial_kernel(const float2* __restrict__ score, const float2* __restrict__ z,
                                                           const float* __restrict__ coef, double* __restrict__ part, int mae,
                                                           unsigned long long seed, unsigned int clip0, size_t per_clip, int B) {
  const int b = blockIdx.y, blk = blockIdx.x;
  const size_t chunk = (per_clip + kRedBlocks - 1) / kRedBlocks;
  const size_t lo = blk * chunk, hi = lo + chunk < per_clip ? lo + chunk : per_clip;
  const float sd = coef[B + b];
  double acc = 0.0;
  for (size_t i = lo + threadIdx.x; i < hi; i += 256) {
    const float2 s = score[static_cast<size_t>(b) * per_clip + i];
    const float2 zz = z ? z[static_cast<size_t>(b) * per_clip + i] : philox_cnormal(seed, 0xfffffffeu, clip0 + b, i);
    const float er = s.x * sd + zz.x, ei = s.y * sd + zz.y;
    const float a2 = er * er + ei * ei;
    acc += mae ? static_cast<double>(sqrtf(a2)) : static_cast<double>(a2);
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[static_cast<size_t>(b) * kRedBlocks + blk] = sh[0];
}

// loss[0] = mean_b 0.5 sum_b ; loss[1 + b] = 0.5 sum_b (per-clip terms, for inspection)
__global__ void dsm_final_kernel(const double* __restrict__ part, float* __restrict__ loss, int B) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float tot = 0.f;
  for (int b = 0; b < B; ++b) {
    double s = 0.0;
    for (int k = 0; k < kRedBlocks; ++k) s += part[static_cast<size_t>(b) * kRedBlocks + k];
    const float lb = 0.5f * static_cast<float>(s);
    loss[1 + b] = lb;
    tot += lb;
  }
  loss[0] = tot / static_cast<float>(B);
}

void launch_perturb(const float2* X0, const float2* Y, const float2* z, const float* coef, float2* xt, unsigned long long seed,
                    unsigned int clip0, int B, size_t per_clip, cudaStream_t st) {
  const size_t n = per_clip * B;
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
  perturb_kernel<<<blocks, 256, 0, st>>>(X0, Y, z, coef, xt, seed, clip0, per_clip, n, B);
}

void launch_dsm_loss(const float2* score, const float2* z, const float* coef, void* scratch, float* loss, int mae,
                     unsigned long long seed, unsigned int clip0, int B, size_t per_clip, cudaStream_t st) {
  double* part = reinterpret_cast<double*>(scratch);
  dsm_partial_kernel<<<dim3(kRedBlocks, B), 256, 0, st>>>(score, z, coef, part, mae, seed, clip0, per_clip, B);
  dsm_final_kernel<<<1, 32, 0, st>>>(part, loss, B);
}

void launch_final_step(const StepArgs& a, cudaStream_t st) {
  const size_t n = a.per_clip * a.B;
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
  final_step_kernel<<<blocks, 256, 0, st>>>(a);
}

__global__ void __launch_bounds__(256) prior_kernel(const float2* __restrict__ Y, const float2* __restrict__ z,
                                                     float2* __restrict__ x0, float std, unsigned long long seed,
                                                     unsigned int clip0, size_t per_clip, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t b = static_cast<uint32_t>(i / per_clip);
    const float2 zz = z ? z[i] : philox_cnormal(seed, 0xffffffffu, clip0 + b, i - static_cast<size_t>(b) * per_clip);
    const float2 y = Y[i];
    x0[i] = make_float2(y.x + zz.x * std, y.y + zz.y * std);
  }
}

void launch_prior(const float2* Y, const float2* z, float2* x0, float std, unsigned long long seed, unsigned int clip0,
                  int B, size_t per_clip, cudaStream_t st) {
  const size_t n = per_clip * B;
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
  prior_kernel<<<blocks, 256, 0, st>>>(Y, z, x0, std, seed, clip0, per_clip, n);
}

__global__ void __launch_bounds__(256) philox_fill_kernel(float2* z, unsigned long long seed, unsigned int step,
                                                           unsigned int clip0, size_t per_clip, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t b = static_cast<uint32_t>(i / per_clip);
    z[i] = philox_cnormal(seed, step, clip0 + b, i - static_cast<size_t>(b) * per_clip);
  }
}

void launch_philox_fill(float2* z, unsigned long long seed, unsigned int step, unsigned int clip0, int B,
                        size_t per_clip, cudaStream_t st) {
  const size_t n = per_clip * B;
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
  philox_fill_kernel<<<blocks, 256, 0, st>>>(z, seed, step, clip0, per_clip, n);
}

}  // namespace use
