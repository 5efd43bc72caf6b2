// STFT / inverse STFT with the spectral (de)compression fused.  n_fft = 1022 = 2*7*73 is not a power of
// two, and the transform runs once per clip (< 0.01 % of the FLOPs), so this is a direct DFT against a
// host-computed (float64-accurate) twiddle table, indexed by (k*n mod n_fft) so no angle is ever large.
// Reference: torch.stft(center=True -> reflect pad n_fft/2, onesided, periodic Hann) and spec_fwd
// (model_wrapper.py:92-96,116-118); spec_back + torch.istft(center=True, length=L): inverse onesided DFT,
// synthesis window, overlap-add, division by the squared-window envelope, trim n_fft/2 (:98-103,120-122).
#include "common.cuh"
#include "kernels.h"

namespace use {

// grid (Tp, B).  smem: frame[n_fft] float, tw[n_fft] float2
__global__ void __launch_bounds__(256) stft_kernel(const float* __restrict__ y, float2* __restrict__ Y,
                                                    const float* __restrict__ window, const float2* __restrict__ twiddle,
                                                    int L, int n_fft, int hop, int T, int Tp, float factor, float exponent) {
  extern __shared__ float sm[];
  float* frame = sm;
  float2* tw = reinterpret_cast<float2*>(sm + ((n_fft + 1) & ~1));
  const int f = blockIdx.x, b = blockIdx.y;
  const int F = n_fft / 2 + 1;
  float2* out = Y + static_cast<size_t>(b) * F * Tp;
  if (f >= T) {  // pad_spec: zero frames up to a multiple of 64
    for (int k = threadIdx.x; k < F; k += blockDim.x) out[static_cast<size_t>(k) * Tp + f] = make_float2(0.f, 0.f);
    return;
  }
  const int half = n_fft / 2;
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    int src = f * hop + n - half;
    if (src < 0) src = -src;
    if (src >= L) src = 2 * (L - 1) - src;
    frame[n] = y[static_cast<size_t>(b) * L + src] * window[n];
    tw[n] = twiddle[n];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < F; k += blockDim.x) {
    float re[2] = {0.f, 0.f}, im[2] = {0.f, 0.f};
    int idx = 0;
    int n = 0;
    for (; n + 1 < n_fft; n += 2) {
      const float2 w0 = tw[idx];
      idx += k; if (idx >= n_fft) idx -= n_fft;
      const float2 w1 = tw[idx];
      idx += k; if (idx >= n_fft) idx -= n_fft;
      const float x0 = frame[n], x1 = frame[n + 1];
      re[0] += x0 * w0.x; im[0] -= x0 * w0.y;
      re[1] += x1 * w1.x; im[1] -= x1 * w1.y;
    }
    if (n < n_fft) {
      const float2 w0 = tw[idx];
      re[0] += frame[n] * w0.x; im[0] -= frame[n] * w0.y;
    }
    const float sr = re[0] + re[1], si = im[0] + im[1];
    // spec_fwd: |S|^e * exp(j angle S) * factor = S * |S|^(e-1) * factor
    const float mag = sqrtf(sr * sr + si * si);
    float g = 0.f;
    if (mag > 0.f) g = (exponent == 0.5f ? rsqrtf(mag) : powf(mag, exponent - 1.0f)) * factor;
    out[static_cast<size_t>(k) * Tp + f] = make_float2(sr * g, si * g);
  }
}

void launch_stft(const float* y, float2* Y, const float* window, const float2* twiddle, int B, int L, int n_fft, int hop,
                 int T, int Tp, float factor, float exponent, cudaStream_t st) {
  dim3 grid(Tp, B);
  const size_t sm = (((n_fft + 1) & ~1) + 2 * n_fft) * sizeof(float);
  stft_kernel<<<grid, 256, sm, st>>>(y, Y, window, twiddle, L, n_fft, hop, T, Tp, factor, exponent);
}

// grid (Tp, B): inverse onesided DFT of one frame, times the synthesis window
__global__ void __launch_bounds__(256) istft_frames_kernel(const float2* __restrict__ X, float* __restrict__ frames,
                                                            const float* __restrict__ window,
                                                            const float2* __restrict__ twiddle, int n_fft, int Tp,
                                                            float factor, float exponent) {
  extern __shared__ float sm[];
  const int F = n_fft / 2 + 1;
  float2* S = reinterpret_cast<float2*>(sm);
  float2* tw = S + F;
  const int f = blockIdx.x, b = blockIdx.y;
  const float2* in = X + static_cast<size_t>(b) * F * Tp;
  const float inv_e = 1.0f / exponent;
  for (int k = threadIdx.x; k < F; k += blockDim.x) {
    float2 v = in[static_cast<size_t>(k) * Tp + f];
    // spec_back: S / factor, then |S|^(1/e) exp(j angle) = S * |S|^(1/e - 1)
    v.x /= factor; v.y /= factor;
    const float mag = sqrtf(v.x * v.x + v.y * v.y);
    const float g = (inv_e == 2.0f) ? mag : (mag > 0.f ? powf(mag, inv_e - 1.0f) : 0.f);
    S[k] = make_float2(v.x * g, v.y * g);
  }
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) tw[n] = twiddle[n];
  __syncthreads();
  const float invn = 1.0f / static_cast<float>(n_fft);
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    // x[n] = (1/N) [ Re S0 + (-1)^n Re S_{N/2} + 2 sum_{k=1}^{N/2-1} (Re S_k cos - Im S_k sin)(2 pi k n / N) ]
    float acc[2] = {0.f, 0.f};
    int idx = n;  // k = 1
    int k = 1;
    for (; k + 1 < F - 1; k += 2) {
      const float2 w0 = tw[idx];
      idx += n; if (idx >= n_fft) idx -= n_fft;
      const float2 w1 = tw[idx];
      idx += n; if (idx >= n_fft) idx -= n_fft;
      acc[0] += S[k].x * w0.x - S[k].y * w0.y;
      acc[1] += S[k + 1].x * w1.x - S[k + 1].y * w1.y;
    }
    for (; k < F - 1; ++k) {
      const float2 w0 = tw[idx];
      idx += n; if (idx >= n_fft) idx -= n_fft;
      acc[0] += S[k].x * w0.x - S[k].y * w0.y;
    }
    const float nyq = (n & 1) ? -S[F - 1].x : S[F - 1].x;
    const float v = (S[0].x + nyq + 2.0f * (acc[0] + acc[1])) * invn;
    frames[(static_cast<size_t>(b) * Tp + f) * n_fft + n] = v * window[n];
  }
}

// overlap-add + envelope normalisation + trim: y[m] = sum_f frames[f][m + N/2 - f hop] / env[m + N/2]
__global__ void __launch_bounds__(256) istft_ola_kernel(const float* __restrict__ frames, const float* __restrict__ env,
                                                         float* __restrict__ y, int L, int n_fft, int hop, int Tp) {
  const int b = blockIdx.y;
  const int half = n_fft / 2;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < L; m += gridDim.x * blockDim.x) {
    const int p = m + half;
    int f_lo = (p - n_fft + 1 + hop - 1) / hop;
    if (p - n_fft + 1 < 0) f_lo = 0;
    int f_hi = p / hop;
    if (f_hi > Tp - 1) f_hi = Tp - 1;
    float acc = 0.f;
    for (int f = f_lo; f <= f_hi; ++f) acc += frames[(static_cast<size_t>(b) * Tp + f) * n_fft + (p - f * hop)];
    y[static_cast<size_t>(b) * L + m] = acc / env[p];
  }
}

void launch_istft(const float2* X, float* frames, float* y, const float* window, const float2* twiddle, const float* env,
                  int B, int L, int n_fft, int hop, int Tp, float factor, float exponent, cudaStream_t st) {
  const int F = n_fft / 2 + 1;
  dim3 grid(Tp, B);
  const size_t sm = (static_cast<size_t>(F) + n_fft) * sizeof(float2);
  istft_frames_kernel<<<grid, 256, sm, st>>>(X, frames, window, twiddle, n_fft, Tp, factor, exponent);
  dim3 grid2((L + 255) / 256, B);
  istft_ola_kernel<<<grid2, 256, 0, st>>>(frames, env, y, L, n_fft, hop, Tp);
}

}  // namespace use
