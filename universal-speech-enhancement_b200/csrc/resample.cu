// Predict-side audio preparation on the GPU: what LoadWavDataset.__getitem__ does on CPU
// workers in the reference (src/data/components/loadwav_dataset.py:90-120) --
//   librosa.resample(y, orig_sr, target_sr, res_type="fft")  == scipy.signal.resample(y, ceil(len * ratio)): rfft, keep
//   the min(n_in, n_out)//2 + 1 low bins (the unpaired middle bin doubled when shrinking / halved when growing),
//   irfft to n_out samples, scale n_out / n_in;
//   y / max|y| * 0.8;
// and the zero padding to the longest clip of collate.pad_to_longest_monaural_inference (collate.py:42-73).
//
// Clip lengths are arbitrary (a 63 997-sample file is legal), so the two DFTs are Bluestein chirp-z transforms: a length-N
// DFT becomes a circular convolution of power-of-two size M >= 2N - 1, computed with radix-2 Stockham FFT passes over
// global memory.  Chirp phases pi n^2 / N are reduced exactly in 64-bit integers (n^2 mod 2N) before the sincospi, so
// the fp32 transform stays at ~1e-6 of the signal's peak for N ~ 1e5.  Once per file, ~100 short launches (~0.3 ms):
// < 0.5 % of the sampling time of the same clip, which is what "keeps up with the GPU sampler" needs.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace use {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// w[n] = exp(-i pi n^2 / N), n < N
__global__ void __launch_bounds__(256) chirp_kernel(float2* __restrict__ w, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const unsigned long long r = (static_cast<unsigned long long>(n) * n) % (2ull * N);
  double s, c;
  sincospi(static_cast<double>(r) / static_cast<double>(N), &s, &c);
  w[n] = make_float2(static_cast<float>(c), static_cast<float>(-s));
}

// the convolution kernel of the chirp-z transform, wrapped to length M: b[j] = conj(w[j]) for |j| < N
__global__ void __launch_bounds__(256) chirp_filter_kernel(const float2* __restrict__ w, float2* __restrict__ b, int N, int M) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  float2 v = make_float2(0.f, 0.f);
  if (j < N) v = make_float2(w[j].x, -w[j].y);
  else if (M - j < N) v = make_float2(w[M - j].x, -w[M - j].y);
  b[j] = v;
}

// one radix-2 Stockham pass over B sequences of length M: n = current sub-transform length, s = stride
__global__ void __launch_bounds__(256) stockham_pass_kernel(const float2* __restrict__ x, float2* __restrict__ y, int M, int n,
                                                             int s, float sign) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = M >> 1;
  if (i >= half) return;
  const size_t base = static_cast<size_t>(blockIdx.y) * M;
  const int m = n >> 1;
  const int p = i / s, q = i - p * s;
  float sn, cs;
  sincospif(sign * 2.0f * static_cast<float>(p) / static_cast<float>(n), &sn, &cs);
  const float2 a = x[base + q + static_cast<size_t>(s) * p];
  const float2 b = x[base + q + static_cast<size_t>(s) * (p + m)];
  y[base + q + static_cast<size_t>(s) * (2 * p)] = make_float2(a.x + b.x, a.y + b.y);
  y[base + q + static_cast<size_t>(s) * (2 * p + 1)] = cmul(make_float2(a.x - b.x, a.y - b.y), make_float2(cs, sn));
}

// in-place-looking FFT of B x M complex values living in buf0 (buf1 = scratch); returns the buffer holding the result
static float2* fft_pow2(float2* buf0, float2* buf1, int B, int M, float sign, cudaStream_t st) {
  float2 *x = buf0, *y = buf1;
  const dim3 grid((M / 2 + 255) / 256, B);
  for (int n = M, s = 1; n > 1; n >>= 1, s <<= 1) {
    stockham_pass_kernel<<<grid, 256, 0, st>>>(x, y, M, n, s, sign);
    std::swap(x, y);
  }
  return x;
}

// a[j] = in[j] * w[j] (j < N), 0 up to M.  REAL: `in` is a real signal; otherwise complex.
template <bool REAL>
__global__ void __launch_bounds__(256) chirp_pre_kernel(const void* __restrict__ in, const float2* __restrict__ w,
                                                         float2* __restrict__ a, int N, int M) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const size_t b = blockIdx.y;
  float2 v = make_float2(0.f, 0.f);
  if (j < N) {
    if constexpr (REAL) {
      const float x = reinterpret_cast<const float*>(in)[b * N + j];
      v = make_float2(x * w[j].x, x * w[j].y);
    } else {
      v = cmul(reinterpret_cast<const float2*>(in)[b * N + j], w[j]);
    }
  }
  a[b * M + j] = v;
}

__global__ void __launch_bounds__(256) pointwise_mul_kernel(float2* __restrict__ a, const float2* __restrict__ bh, int M) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const size_t i = static_cast<size_t>(blockIdx.y) * M + j;
  a[i] = cmul(a[i], bh[j]);
}

// Forward half done: X[k] = w[k] c[k] / M for the m2 kept bins, then straight into the Hermitian spectrum the inverse
// transform consumes, conjugated (IDFT(Y) = conj(DFT(conj Y))):  yc[k] = conj(Y[k]), Y = irfft's full spectrum of
// X[:m2] * f (f = n_out / n_in, the unpaired bin m/2 doubled when shrinking, halved when growing)
__global__ void __launch_bounds__(256) spectrum_resize_kernel(const float2* __restrict__ c, const float2* __restrict__ w_in,
                                                               float2* __restrict__ yc, int n_in, int n_out, int M1) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_out) return;
  const size_t b = blockIdx.y;
  const int m = min(n_in, n_out), m2 = m / 2 + 1;
  const int kk = k <= n_out / 2 ? k : n_out - k;  // the one-sided bin this full-spectrum bin mirrors
  float2 v = make_float2(0.f, 0.f);
  if (kk < m2) {
    const float2 X = cmul(c[b * M1 + kk], w_in[kk]);
    float f = 1.0f / static_cast<float>(M1);
    if ((m % 2 == 0) && n_out != n_in && kk == m / 2) f *= (n_out < n_in) ? 2.0f : 0.5f;
    v = make_float2(X.x * f, X.y * f);
    if (kk == 0 || (n_out % 2 == 0 && kk == n_out / 2)) v.y = 0.f;  // irfft ignores the imaginary part of DC / Nyquist
    if (k != kk) v.y = -v.y;                                          // negative frequencies: conj(Z[kk])
  }
  yc[b * n_out + k] = make_float2(v.x, -v.y);  // conjugated for the forward-transform trick
}

// y[j] = Re(w_out[j] c[j]) / (M2 * n_in)
__global__ void __launch_bounds__(256) chirp_post_real_kernel(const float2* __restrict__ c, const float2* __restrict__ w,
                                                               float* __restrict__ y, int N, int M, float scale, int y_stride) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const size_t b = blockIdx.y;
  const float2 v = cmul(c[b * M + j], w[j]);
  y[b * y_stride + j] = v.x * scale;
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

size_t resample_workspace_bytes(int B, int n_in, int n_out) {
  const size_t M1 = next_pow2(2 * n_in - 1), M2 = next_pow2(2 * n_out - 1);
  // chirps w_in [n_in], w_out [n_out]; filter spectra [M1], [M2] (+ one scratch each); two work buffers of B * max(M1, M2);
  // the resized spectrum B * n_out
  const size_t mm = std::max(M1, M2);
  return sizeof(float2) * (n_in + n_out + 2 * M1 + 2 * M2 + 2 * B * mm + static_cast<size_t>(B) * n_out) + 4096;
}

void launch_resample_fft(const float* x, float* y, int B, int n_in, int n_out, int y_stride, void* work, cudaStream_t st) {
  const int M1 = next_pow2(2 * n_in - 1), M2 = next_pow2(2 * n_out - 1);
  const size_t mm = std::max(M1, M2);
  float2* p = reinterpret_cast<float2*>(work);
  float2* w_in = p; p += n_in;
  float2* w_out = p; p += n_out;
  float2* f1 = p; p += M1;
  float2* f1s = p; p += M1;
  float2* f2 = p; p += M2;
  float2* f2s = p; p += M2;
  float2* a0 = p; p += static_cast<size_t>(B) * mm;
  float2* a1 = p; p += static_cast<size_t>(B) * mm;
  float2* yc = p;
  auto blocks = [](int n) { return (n + 255) / 256; };
  // chirps and the spectra of the two convolution kernels
  chirp_kernel<<<blocks(n_in), 256, 0, st>>>(w_in, n_in);
  chirp_kernel<<<blocks(n_out), 256, 0, st>>>(w_out, n_out);
  chirp_filter_kernel<<<blocks(M1), 256, 0, st>>>(w_in, f1, n_in, M1);
  chirp_filter_kernel<<<blocks(M2), 256, 0, st>>>(w_out, f2, n_out, M2);
  const float2* F1 = fft_pow2(f1, f1s, 1, M1, -1.f, st);
  const float2* F2 = fft_pow2(f2, f2s, 1, M2, -1.f, st);
  // ---- forward DFT of the real clips (length n_in) ----
  chirp_pre_kernel<true><<<dim3(blocks(M1), B), 256, 0, st>>>(x, w_in, a0, n_in, M1);
  float2* A = fft_pow2(a0, a1, B, M1, -1.f, st);
  pointwise_mul_kernel<<<dim3(blocks(M1), B), 256, 0, st>>>(A, F1, M1);
  float2* C1 = fft_pow2(A, A == a0 ? a1 : a0, B, M1, +1.f, st);
  // ---- keep the low bins, rebuild the Hermitian spectrum of length n_out (conjugated) ----
  spectrum_resize_kernel<<<dim3(blocks(n_out), B), 256, 0, st>>>(C1, w_in, yc, n_in, n_out, M1);
  // ---- inverse DFT (length n_out) as a forward transform of the conjugate ----
  chirp_pre_kernel<false><<<dim3(blocks(M2), B), 256, 0, st>>>(yc, w_out, a0, n_out, M2);
  A = fft_pow2(a0, a1, B, M2, -1.f, st);
  pointwise_mul_kernel<<<dim3(blocks(M2), B), 256, 0, st>>>(A, F2, M2);
  float2* C2 = fft_pow2(A, A == a0 ? a1 : a0, B, M2, +1.f, st);
  chirp_post_real_kernel<<<dim3(blocks(n_out), B), 256, 0, st>>>(C2, w_out, y, n_out, M2,
                                                                 1.0f / (static_cast<float>(M2) * static_cast<float>(n_in)), y_stride);
}

// ---- peak normalisation (loadwav_dataset.py:99-100) + zero padding (collate.py:59) ------------------------------------
// peaks[b] = max |y[b][0 .. len_b)| as float bits (non-negative floats order like unsigned integers: atomicMax is exact and
// order-independent); zero `peaks` first
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ y, const int* __restrict__ lengths, int stride,
                                                      unsigned int* __restrict__ peaks) {
  const int b = blockIdx.y;
  const int len = lengths[b];
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(y[static_cast<size_t>(b) * stride + i]));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(peaks + b, __float_as_uint(m));
}

// y[b][i] = i < len_b ? y[b][i] / peak_b * target : 0   (an all-zero clip stays zero)
__global__ void __launch_bounds__(256) scale_pad_kernel(float* __restrict__ y, const int* __restrict__ lengths, int stride,
                                                         const unsigned int* __restrict__ peaks, float target) {
  const int b = blockIdx.y;
  const int len = lengths[b];
  const float peak = __uint_as_float(peaks[b]);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < stride; i += gridDim.x * blockDim.x) {
    const size_t o = static_cast<size_t>(b) * stride + i;
    float v = 0.f;
    if (i < len) v = (target > 0.f && peak > 0.f) ? y[o] / peak * target : y[o];
    y[o] = v;
  }
}

void launch_peak_normalize_pad(float* y, const int* lengths, int B, int stride, float target, unsigned int* peaks,
                               cudaStream_t st) {
  cudaMemsetAsync(peaks, 0, sizeof(unsigned int) * B, st);
  const dim3 grid(std::min((stride + 255) / 256, 64), B);
  if (target > 0.f) absmax_kernel<<<grid, 256, 0, st>>>(y, lengths, stride, peaks);
  scale_pad_kernel<<<grid, 256, 0, st>>>(y, lengths, stride, peaks, target);
}

}  // namespace use
