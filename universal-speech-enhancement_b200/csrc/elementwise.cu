// Bandwidth-bound kernels of the score network: GroupNorm statistics, normalise + SiLU (+ FIR resample),
// Combine, the 4-channel input / pyramid convolutions and the FIR of the input pyramid.
// Reference semantics: nn.GroupNorm(min(C/4,32), eps=1e-6) + SiLU (layerspp.py:255-271,283,304),
// upsample_2d / downsample_2d with k=[1,3,3,1] (up_or_down_sampling.py:202-264), Combine "sum"
// (layerspp.py:50-55), input conv / pyramid convs (ncsnpp.py:214,377-381,440-461).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace use {

#define DISPATCH_DT(dt, ...)                \
  do {                                      \
    if ((dt) == kBF16) {                    \
      using T = __nv_bfloat16;              \
      __VA_ARGS__                           \
    } else {                                \
      using T = float;                      \
      __VA_ARGS__                           \
    }                                       \
  } while (0)

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics of a tensor that no conv epilogue produced (Combine outputs): per sample and per CHANNEL
// sum and sum of squares.  Channels (not groups) so that a GroupNorm over a channel concat whose group boundary
// straddles the two sources (384 = 256 + 128 channels, 12 per group) can be assembled from per-tensor partials.
// Every block reduces a fixed pixel range in a fixed order and adds its partial with 64-bit fixed-point integer
// atomics (common.cuh): bit-reproducible and independent of the batch size.  (Any run-to-run ulp difference would be
// amplified to the bf16 / TF32 rounding-noise floor within a few layers, which would break "a clip sampled alone ==
// the same clip sampled in a batch".)
// ------------------------------------------------------------------------------------------------
constexpr int kGnPixPerBlock = 2048;

template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, long long* __restrict__ stats, int HW, int C) {
  constexpr int V = Vec<T>::N;
  extern __shared__ float sred[];  // [rows][C][2]
  const int b = blockIdx.y;
  const int vp = C / V;  // vectors per pixel
  const int p0 = blockIdx.x * kGnPixPerBlock;
  const int p1 = min(HW, p0 + kGnPixPerBlock);
  const T* base = x + (static_cast<size_t>(b) * HW) * C;
  const int lanes = min(vp, static_cast<int>(blockDim.x));
  const int rows = max(1, static_cast<int>(blockDim.x) / vp);
  const int nact = rows * lanes;
  for (int v0 = 0; v0 < vp; v0 += blockDim.x) {  // vp > blockDim only for very wide tensors
    const int v = v0 + (threadIdx.x % lanes);
    const int r = threadIdx.x / lanes;
    float s[V], ss[V];
#pragma unroll
    for (int i = 0; i < V; ++i) s[i] = ss[i] = 0.f;
    if (threadIdx.x < nact && v < vp) {
      for (int p = p0 + r; p < p1; p += rows) {
        float f[V];
        Vec<T>::load(base + static_cast<size_t>(p) * C + v * V, f);
#pragma unroll
        for (int i = 0; i < V; ++i) { s[i] += f[i]; ss[i] += f[i] * f[i]; }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) {
        sred[(r * C + v * V + i) * 2] = s[i];
        sred[(r * C + v * V + i) * 2 + 1] = ss[i];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double a = 0.0, q = 0.0;
    for (int r = 0; r < rows; ++r) {
      a += static_cast<double>(sred[(r * C + c) * 2]);
      q += static_cast<double>(sred[(r * C + c) * 2 + 1]);
    }
    stat_atomic_add(stats + (static_cast<size_t>(b) * C + c) * 2, a, q);
  }
}

void launch_gn_stats(int dt, const void* x, long long* stats, int B, int HW, int C, cudaStream_t st) {
  dim3 grid((HW + kGnPixPerBlock - 1) / kGnPixPerBlock, B);
  DISPATCH_DT(dt, {
    constexpr int V = Vec<T>::N;
    const int vp = C / V;
    const int rows = vp >= 256 ? 1 : 256 / vp;
    gn_stats_kernel<T><<<grid, 256, static_cast<size_t>(rows) * C * 2 * sizeof(float), st>>>((const T*)x, stats, HW, C);
  });
}

// ------------------------------------------------------------------------------------------------
// normalise + affine + SiLU (+ FIR x2 up / down) over a (possibly concatenated) input.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct GnSrcT {
  const T* x;
  const long long* stats;
  int C;
};

// Scale / shift table of a GroupNorm over cat[s0, s1] for the convolution kernel's fused operand path: one block per
// sample, identical arithmetic to the prologue of gn_apply_kernel (the two paths give bit-identical operands).
__global__ void __launch_bounds__(256) gn_affine_kernel(const long long* __restrict__ st0, int C0,
                                                         const long long* __restrict__ st1, int C1,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float eps, int HW, float* __restrict__ aff) {
  // PDL: this kernel sits between two convolutions; it lets the consumer convolution start its set-up now, and (being
  // launched with the attribute itself) is already resident when the producer convolution finishes
  pdl_launch_dependents();
  pdl_wait();
  const int Ct = C0 + C1;
  const int G = min(Ct / 4, 32);
  const int cpg = Ct / G;
  const int b = blockIdx.x;
  // Two phases: every thread first fetches ONE channel's statistics (all loads of the block in flight at once), then sums
  // its group from shared memory.  The one-phase loop walked a group's cpg = 4 - 16 channels with dependent global loads:
  // one L2 round trip per channel, ~3 - 11 us of a kernel that sits between two convolutions 98 times per evaluation.
  // Same products in the same order: bit-identical.
  extern __shared__ double gst[];  // [Ct][2]
  for (int c = threadIdx.x; c < Ct; c += blockDim.x) {
    const longlong2 st = __ldg(reinterpret_cast<const longlong2*>(
        (c < C0) ? st0 + (static_cast<size_t>(b) * C0 + c) * 2 : st1 + (static_cast<size_t>(b) * C1 + (c - C0)) * 2));
    gst[2 * c] = static_cast<double>(st.x) * (1.0 / kStatSumScale);
    gst[2 * c + 1] = static_cast<double>(st.y) * (1.0 / kStatSqScale);
  }
  __syncthreads();
  const double inv_cnt = 1.0 / (static_cast<double>(HW) * cpg);
  for (int c = threadIdx.x; c < Ct; c += blockDim.x) {
    const int g = c / cpg;
    double sum = 0.0, sq = 0.0;
    for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {
      sum += gst[2 * cc];
      sq += gst[2 * cc + 1];
    }
    const double mean = sum * inv_cnt;
    double var = sq * inv_cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = rsqrtf(static_cast<float>(var) + eps);
    const float sc = gamma[c] * rstd;
    aff[(static_cast<size_t>(b) * 2) * Ct + c] = sc;
    aff[(static_cast<size_t>(b) * 2 + 1) * Ct + c] = beta[c] - static_cast<float>(mean) * sc;
  }
}

bool pdl_enabled(int B) {
  static const int force = [] {
    const char* v = getenv("USE_B200_PDL");
    return !v ? -1 : (v[0] == '1' ? 1 : 0);
  }();
  static const int maxb = getenv("USE_B200_PDL_MAXB") ? atoi(getenv("USE_B200_PDL_MAXB")) : 2;
  return force >= 0 ? force == 1 : B <= maxb;
}

void launch_gn_affine(GnSrc s0, GnSrc s1, const float* gamma, const float* beta, float eps, int HW, float* aff, int B,
                      cudaStream_t st) {
  const bool pdl = pdl_enabled(B);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(B);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = static_cast<size_t>(s0.C + s1.C) * 2 * sizeof(double);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, gn_affine_kernel, s0.stats, s0.C, s1.stats, s1.C, gamma, beta, eps, HW, aff);
}

// One thread owns a fixed 16-byte channel vector (scale / shift live in registers) and walks over pixels;
// consecutive threads cover consecutive vectors of the same pixel, so every warp access is a contiguous run.
constexpr int kGnMaxWorkPerBlock = 256;  // work items (pixels / 2x2 blocks) per block; shrunk for small tensors

template <typename T, int FIR>
__global__ void __launch_bounds__(256) gn_apply_kernel(GnSrcT<T> s0, GnSrcT<T> s1, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int do_silu, int as_operand,
                                                        T* __restrict__ out_act, T* __restrict__ out_raw, int Hin, int Win,
                                                        int work_per_block) {
  constexpr int V = Vec<T>::N;
  extern __shared__ double gst_[];  // statistics [Ct][2] as doubles, then scale[Ct], shift[Ct]
  const int Ct = s0.C + s1.C;
  float* saff = reinterpret_cast<float*>(gst_ + 2 * Ct);
  const int G = min(Ct / 4, 32);
  const int cpg = Ct / G;
  const int b = blockIdx.y;
  const double inv_cnt = 1.0 / (static_cast<double>(Hin) * Win * cpg);
  // (two phases as in gn_affine_kernel: one statistics load per thread, group sums from shared memory)
  for (int c = threadIdx.x; c < Ct; c += blockDim.x) {
    const longlong2 st = __ldg(reinterpret_cast<const longlong2*>(
        (c < s0.C) ? s0.stats + (static_cast<size_t>(b) * s0.C + c) * 2
                   : s1.stats + (static_cast<size_t>(b) * s1.C + (c - s0.C)) * 2));
    gst_[2 * c] = static_cast<double>(st.x) * (1.0 / kStatSumScale);
    gst_[2 * c + 1] = static_cast<double>(st.y) * (1.0 / kStatSqScale);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Ct; c += blockDim.x) {
    const int g = c / cpg;
    double sum = 0.0, sq = 0.0;
    for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {  // channel index in the concatenation
      sum += gst_[2 * cc];
      sq += gst_[2 * cc + 1];
    }
    const double mean = sum * inv_cnt;
    double var = sq * inv_cnt - mean * mean;  // biased variance, like nn.GroupNorm
    if (var < 0.0) var = 0.0;
    const float rstd = rsqrtf(static_cast<float>(var) + eps);
    const float sc = gamma[c] * rstd;
    saff[c] = sc;
    saff[Ct + c] = beta[c] - static_cast<float>(mean) * sc;
  }
  __syncthreads();

  const int Hout = FIR == 1 ? Hin / 2 : (FIR == 2 ? Hin * 2 : Hin);
  const int Wout = FIR == 1 ? Win / 2 : (FIR == 2 ? Win * 2 : Win);
  const int vpp = Ct / V;
  const int rows = blockDim.x / vpp;  // blockDim.x is a multiple of vpp (host guarantees it)
  const int cv = threadIdx.x % vpp, prow = threadIdx.x / vpp;
  const int c = cv * V;
  const bool first = c < s0.C;
  const int Cs = first ? s0.C : s1.C;
  const T* sb = (first ? s0.x + c : s1.x + (c - s0.C)) + static_cast<size_t>(b) * Hin * Win * Cs;
  float sc[V], sh[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { sc[j] = saff[c + j]; sh[j] = saff[Ct + c + j]; }
  const int npix = Hout * Wout;
  // work items: output pixels (FIR 0), input pixels (FIR 2: one 2x2 output block each), 2x2 output blocks (FIR 1)
  const int nwork = FIR == 0 ? npix : (FIR == 2 ? Hin * Win : ((Hout + 1) / 2) * ((Wout + 1) / 2));
  const int p0 = blockIdx.x * work_per_block;
  const int p1 = min(nwork, p0 + work_per_block);
  T* oa = out_act + static_cast<size_t>(b) * npix * Ct + c;
  T* orw = (FIR != 0 && out_raw != nullptr) ? out_raw + static_cast<size_t>(b) * npix * Ct + c : nullptr;

  if constexpr (FIR == 0) {
    constexpr int U = 4;  // independent 16-byte loads in flight per thread
    const int n = (p1 - p0 - prow + rows - 1) / rows;  // pixels of this thread: p0 + prow + k * rows, k < n
    const T* src = sb + static_cast<size_t>(p0 + prow) * Cs;
    T* dst = oa + static_cast<size_t>(p0 + prow) * Ct;
    const size_t sstep = static_cast<size_t>(rows) * Cs, dstep = static_cast<size_t>(rows) * Ct;
    int k = 0;
    for (; k + U <= n; k += U) {
      float f[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) Vec<T>::load(src + (k + u) * sstep, f[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float nn = fmaf(f[u][j], sc[j], sh[j]);
          f[u][j] = do_silu ? silu_act<T>(nn) : nn;
        }
        if (as_operand) Vec<T>::store_operand(dst + (k + u) * dstep, f[u]);
        else Vec<T>::store(dst + (k + u) * dstep, f[u]);
      }
    }
    for (; k < n; ++k) {
      float f[V];
      Vec<T>::load(src + k * sstep, f);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float nn = fmaf(f[j], sc[j], sh[j]);
        f[j] = do_silu ? silu_act<T>(nn) : nn;
      }
      if (as_operand) Vec<T>::store_operand(dst + k * dstep, f);
      else Vec<T>::store(dst + k * dstep, f);
    }
  } else if constexpr (FIR == 2) {
    // x2 upsample: each thread produces the 2x2 output block of input pixel (m, n) from its 3x3 neighbourhood
    // (9 normalise+SiLU evaluations per 4 outputs instead of 16).  Per axis:
    //   out[2m] = (in[m-1] + 3 in[m]) / 4 ; out[2m+1] = (3 in[m] + in[m+1]) / 4, zero outside.
    for (int p = p0 + prow; p < p1; p += rows) {   // p indexes INPUT pixels here (grid sized by the host accordingly)
      const int m = p / Win, n = p - m * Win;
      float ea[2][V], oa2[2][V], er[2][V], orr[2][V];  // [0]: output column 2n, [1]: output column 2n+1; e = row 2m, o = row 2m+1
#pragma unroll
      for (int j = 0; j < V; ++j) ea[0][j] = ea[1][j] = oa2[0][j] = oa2[1][j] = er[0][j] = er[1][j] = orr[0][j] = orr[1][j] = 0.f;
#pragma unroll
      for (int dc = -1; dc <= 1; ++dc) {
        const int xx = n + dc;
        if (xx < 0 || xx >= Win) continue;
        float ve[V], vo[V], re[V], ro[V];  // vertical combinations for this input column
#pragma unroll
        for (int j = 0; j < V; ++j) ve[j] = vo[j] = re[j] = ro[j] = 0.f;
#pragma unroll
        for (int dr = -1; dr <= 1; ++dr) {
          const int yy = m + dr;
          if (yy < 0 || yy >= Hin) continue;
          float f[V];
          Vec<T>::load(sb + (static_cast<size_t>(yy) * Win + xx) * Cs, f);
          const float we = dr == -1 ? 0.25f : (dr == 0 ? 0.75f : 0.f);
          const float wo = dr == -1 ? 0.f : (dr == 0 ? 0.75f : 0.25f);
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const float nn = fmaf(f[j], sc[j], sh[j]);
            const float a = do_silu ? silu_act<T>(nn) : nn;
            ve[j] += we * a; vo[j] += wo * a;
            re[j] += we * f[j]; ro[j] += wo * f[j];
          }
        }
        const float w0 = dc == -1 ? 0.25f : (dc == 0 ? 0.75f : 0.f);   // weight into output column 2n
        const float w1 = dc == -1 ? 0.f : (dc == 0 ? 0.75f : 0.25f);   // weight into output column 2n+1
#pragma unroll
        for (int j = 0; j < V; ++j) {
          ea[0][j] += w0 * ve[j]; ea[1][j] += w1 * ve[j];
          oa2[0][j] += w0 * vo[j]; oa2[1][j] += w1 * vo[j];
          er[0][j] += w0 * re[j]; er[1][j] += w1 * re[j];
          orr[0][j] += w0 * ro[j]; orr[1][j] += w1 * ro[j];
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const size_t o = (static_cast<size_t>(2 * m + (q >> 1)) * Wout + 2 * n + (q & 1)) * Ct;
        const float(&va)[V] = (q >> 1) ? oa2[q & 1] : ea[q & 1];
        const float(&vr)[V] = (q >> 1) ? orr[q & 1] : er[q & 1];
        if (as_operand) Vec<T>::store_operand(oa + o, va);
        else Vec<T>::store(oa + o, va);
        if (orw != nullptr) {
          if (as_operand) Vec<T>::store_operand(orw + o, vr);
          else Vec<T>::store(orw + o, vr);
        }
      }
    }
  } else {
    // x2 downsample: out[i][j] = sum_{a,b<4} k[a]k[b] in[2i+a-1][2j+b-1], k = [1,3,3,1]/8, zero outside.
    // Each thread produces a 2x2 output block from a 6x6 input window, separably (36 normalise+SiLU evaluations
    // per 4 outputs instead of 64).  p indexes 2x2 output blocks.
    const int bw2 = (Wout + 1) / 2;
    const float k1[4] = {0.125f, 0.375f, 0.375f, 0.125f};
    for (int p = p0 + prow; p < p1; p += rows) {
      const int I = p / bw2, J = p - I * bw2;
      float acc[4][V], raw[4][V];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[q][j] = raw[q][j] = 0.f;
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        const int ix = 4 * J - 1 + cc;
        if (ix < 0 || ix >= Win) continue;
        float va[2][V], vr[2][V];  // vertical FIR for output rows 2I and 2I+1 at this input column
#pragma unroll
        for (int j = 0; j < V; ++j) va[0][j] = va[1][j] = vr[0][j] = vr[1][j] = 0.f;
#pragma unroll
        for (int rr = 0; rr < 6; ++rr) {
          const int iy = 4 * I - 1 + rr;
          if (iy < 0 || iy >= Hin) continue;
          float f[V];
          Vec<T>::load(sb + (static_cast<size_t>(iy) * Win + ix) * Cs, f);
          const float w0 = rr < 4 ? k1[rr] : 0.f;       // tap of output row 2I
          const float w1 = rr >= 2 ? k1[rr - 2] : 0.f;  // tap of output row 2I+1
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const float nn = fmaf(f[j], sc[j], sh[j]);
            const float a = do_silu ? silu_act<T>(nn) : nn;
            va[0][j] += w0 * a; va[1][j] += w1 * a;
            vr[0][j] += w0 * f[j]; vr[1][j] += w1 * f[j];
          }
        }
        const float h0 = cc < 4 ? k1[cc] : 0.f;
        const float h1 = cc >= 2 ? k1[cc - 2] : 0.f;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          acc[0][j] += h0 * va[0][j]; acc[1][j] += h1 * va[0][j];
          acc[2][j] += h0 * va[1][j]; acc[3][j] += h1 * va[1][j];
          raw[0][j] += h0 * vr[0][j]; raw[1][j] += h1 * vr[0][j];
          raw[2][j] += h0 * vr[1][j]; raw[3][j] += h1 * vr[1][j];
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int oy = 2 * I + (q >> 1), ox = 2 * J + (q & 1);
        if (oy >= Hout || ox >= Wout) continue;
        const size_t o = (static_cast<size_t>(oy) * Wout + ox) * Ct;
        if (as_operand) Vec<T>::store_operand(oa + o, acc[q]);
        else Vec<T>::store(oa + o, acc[q]);
        if (orw != nullptr) {
          if (as_operand) Vec<T>::store_operand(orw + o, raw[q]);
          else Vec<T>::store(orw + o, raw[q]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Shared-memory tiled normalise + SiLU + FIR x2 (single source): every input element is loaded and activated ONCE
// per block, then the 16-tap (down) or 4-tap (up) filter runs out of shared memory for both the activated and the
// raw tensor.  Each thread produces a 2x2 output block for 4 channels, separably, so a filter tap costs 9 (down) /
// 2.25 (up) shared-memory reads per output instead of 16 / 4.  The window is staged in the act dtype (bf16 mode: the
// same rounding the activated tensor would get if it were materialised, and half the shared-memory bytes).
// Block = one output tile x CH channels x one sample.  Down: 8x8 outputs from an 18x18 window; up: 16x16 from 10x10.
// ------------------------------------------------------------------------------------------------
template <typename T> struct FirCh { static constexpr int value = DT<T>::kIsBf16 ? 64 : 32; };

template <typename T> __device__ __forceinline__ float4 lds4(const T* p);
template <> __device__ __forceinline__ float4 lds4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 lds4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u), __uint_as_float(t.y << 16),
                     __uint_as_float(t.y & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ void sts_vec(T* p, const float (&v)[Vec<T>::N]);
template <> __device__ __forceinline__ void sts_vec<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void sts_vec<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  Vec<__nv_bfloat16>::store(p, v);
}
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
}

template <typename T, int FIR>
__global__ void __launch_bounds__(256) gn_apply_fir_tiled_kernel(GnSrcT<T> s0, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float eps, int do_silu,
                                                                  int as_operand, T* __restrict__ out_act,
                                                                  T* __restrict__ out_raw, int Hin, int Win,
                                                                  const float* __restrict__ aff) {
  constexpr int V = Vec<T>::N;
  constexpr int CH = FirCh<T>::value;
  constexpr int TO = FIR == 1 ? 8 : 16;        // output tile edge
  constexpr int WIN = FIR == 1 ? 18 : 10;      // input window edge
  constexpr int VPC = CH / V;                  // 16-byte vectors per pixel of the channel chunk
  extern __shared__ __align__(16) unsigned char smraw[];
  T* sa = reinterpret_cast<T*>(smraw);         // activated  [WIN*WIN][CH]
  T* sr = sa + WIN * WIN * CH;                 // raw        [WIN*WIN][CH]
  float* saff = reinterpret_cast<float*>(sr + WIN * WIN * CH);  // scale[CH], shift[CH]
  const int C = s0.C;
  const int G = min(C / 4, 32), cpg = C / G;
  const int b = blockIdx.z, c0 = blockIdx.y * CH;
  const int Hout = FIR == 1 ? Hin / 2 : Hin * 2, Wout = FIR == 1 ? Win / 2 : Win * 2;
  const int tiles_x = (Wout + TO - 1) / TO;
  const int oy0 = (blockIdx.x / tiles_x) * TO, ox0 = (blockIdx.x % tiles_x) * TO;
  if (threadIdx.x < CH && aff != nullptr) {
    const int c = c0 + threadIdx.x;
    saff[threadIdx.x] = __ldg(aff + (static_cast<size_t>(b) * 2) * C + c);
    saff[CH + threadIdx.x] = __ldg(aff + (static_cast<size_t>(b) * 2 + 1) * C + c);
  } else if (threadIdx.x < CH) {
    const int c = c0 + threadIdx.x, g = c / cpg;
    const double inv_cnt = 1.0 / (static_cast<double>(Hin) * Win * cpg);
    double sum = 0.0, sq = 0.0;
    for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {
      const longlong2 st = __ldg(reinterpret_cast<const longlong2*>(s0.stats + (static_cast<size_t>(b) * C + cc) * 2));
      sum += static_cast<double>(st.x) * (1.0 / kStatSumScale);
      sq += static_cast<double>(st.y) * (1.0 / kStatSqScale);
    }
    const double mean = sum * inv_cnt;
    double var = sq * inv_cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = rsqrtf(static_cast<float>(var) + eps);
    const float sc = gamma[c] * rstd;
    saff[threadIdx.x] = sc;
    saff[CH + threadIdx.x] = beta[c] - static_cast<float>(mean) * sc;
  }
  __syncthreads();
  // ---- load (all requests in flight first) + activate the input window; zero outside the image: the FIR pads the
  //      ACTIVATED tensor with zeros ----
  const int iy0 = FIR == 1 ? 2 * oy0 - 1 : oy0 / 2 - 1, ix0 = FIR == 1 ? 2 * ox0 - 1 : ox0 / 2 - 1;
  const T* src = s0.x + static_cast<size_t>(b) * Hin * Win * C + c0;
  constexpr int NIT = (WIN * WIN * VPC + 255) / 256;
  uint4 raw4[NIT];
#pragma unroll
  for (int k = 0; k < NIT; ++k) {
    const int it = threadIdx.x + k * 256;
    const int v = it % VPC, pw = it / VPC;
    const int iy = iy0 + pw / WIN, ix = ix0 + pw % WIN;
    const bool in = it < WIN * WIN * VPC && iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
    raw4[k] = in ? __ldg(reinterpret_cast<const uint4*>(src + (static_cast<size_t>(iy) * Win + ix) * C + v * V))
                 : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int k = 0; k < NIT; ++k) {
    const int it = threadIdx.x + k * 256;
    if (it >= WIN * WIN * VPC) break;
    const int v = it % VPC, pw = it / VPC;
    const int iy = iy0 + pw / WIN, ix = ix0 + pw % WIN;
    const bool in = iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
    float f[V], a[V];
    Vec<T>::unpack(raw4[k], f);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float n = fmaf(f[j], saff[v * V + j], saff[CH + v * V + j]);
      a[j] = in ? (do_silu ? silu_act<T>(n) : n) : 0.f;
    }
    sts_vec<T>(sa + pw * CH + v * V, a);
    sts_vec<T>(sr + pw * CH + v * V, f);
  }
  __syncthreads();
  // ---- FIR out of shared memory: one item = (2x2 output block, 4-channel quad) ----
  constexpr int NB = TO / 2;  // 2x2 blocks per tile edge
  for (int it = threadIdx.x; it < NB * NB * (CH / 4); it += blockDim.x) {
    const int q = it % (CH / 4), bl = it / (CH / 4);
    const int by = bl / NB, bx = bl % NB;
    float4 acc[2][2], raw[2][2];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) acc[dy][dx] = raw[dy][dx] = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (FIR == 1) {
      // outputs (2by+dy, 2bx+dx) read window rows 4by + 2dy + a, a < 4: a 6x6 window, filtered separably
      const float k1[4] = {0.125f, 0.375f, 0.375f, 0.125f};
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        float4 va[2], vr[2];
        va[0] = va[1] = vr[0] = vr[1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int rr = 0; rr < 6; ++rr) {
          const int pw = (4 * by + rr) * WIN + 4 * bx + cc;
          const float4 xa = lds4<T>(sa + pw * CH + q * 4), xr = lds4<T>(sr + pw * CH + q * 4);
          if (rr < 4) { fma4(va[0], k1[rr], xa); fma4(vr[0], k1[rr], xr); }
          if (rr >= 2) { fma4(va[1], k1[rr - 2], xa); fma4(vr[1], k1[rr - 2], xr); }
        }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          if (cc < 4) { fma4(acc[dy][0], k1[cc], va[dy]); fma4(raw[dy][0], k1[cc], vr[dy]); }
          if (cc >= 2) { fma4(acc[dy][1], k1[cc - 2], va[dy]); fma4(raw[dy][1], k1[cc - 2], vr[dy]); }
        }
      }
    } else {
      // input pixel (by, bx) of the tile = window (by+1, bx+1); out[2m] = (in[m-1] + 3 in[m])/4, out[2m+1] = (3 in[m] + in[m+1])/4
      const float we[3] = {0.25f, 0.75f, 0.f}, wo[3] = {0.f, 0.75f, 0.25f};
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        float4 va[2], vr[2];
        va[0] = va[1] = vr[0] = vr[1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
          const int pw = (by + rr) * WIN + bx + cc;
          const float4 xa = lds4<T>(sa + pw * CH + q * 4), xr = lds4<T>(sr + pw * CH + q * 4);
          if (rr < 2) { fma4(va[0], we[rr], xa); fma4(vr[0], we[rr], xr); }
          if (rr >= 1) { fma4(va[1], wo[rr], xa); fma4(vr[1], wo[rr], xr); }
        }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          if (cc < 2) { fma4(acc[dy][0], we[cc], va[dy]); fma4(raw[dy][0], we[cc], vr[dy]); }
          if (cc >= 1) { fma4(acc[dy][1], wo[cc], va[dy]); fma4(raw[dy][1], wo[cc], vr[dy]); }
        }
      }
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int oy = oy0 + 2 * by + dy, ox = ox0 + 2 * bx + dx;
        if (oy >= Hout || ox >= Wout) continue;
        const size_t o = ((static_cast<size_t>(b) * Hout + oy) * Wout + ox) * C + c0 + q * 4;
        auto put = [&](T* dst, const float4& v4) {
          if constexpr (DT<T>::kIsBf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v4.x, v4.y), hi = __floats2bfloat162_rn(v4.z, v4.w);
            *reinterpret_cast<uint2*>(dst + o) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
          } else {
            *reinterpret_cast<float4*>(dst + o) =
                as_operand ? make_float4(round_tf32(v4.x), round_tf32(v4.y), round_tf32(v4.z), round_tf32(v4.w)) : v4;
          }
        };
        put(out_act, acc[dy][dx]);
        if (out_raw != nullptr) put(out_raw, raw[dy][dx]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// normalise + SiLU + FIR x2 DOWN (single source), second generation.  The first tiled kernel spent 41 instructions per
// input element (ncu: 35 % of them useful math, half of the threads idle in the filter phase in fp32, 2 blocks per SM).
// Here: 8 x 8 outputs from an 18 x 18 window for a 64-byte channel slice (16 fp32 / 32 bf16 channels): 41.5 KB of
// shared memory per block (4-5 blocks per SM); every input element is loaded and activated once; the filter phase gives
// every thread one output pixel x 4 channels with all sixteen taps at compile-time offsets from one base pointer.
// ------------------------------------------------------------------------------------------------
template <typename T> struct FirDownCh { static constexpr int value = 64 / sizeof(T); };

template <typename T>
__global__ void __launch_bounds__(256, 4) gn_fir_down_kernel(GnSrcT<T> s0, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, int do_silu,
                                                              int as_operand, T* __restrict__ out_act, T* __restrict__ out_raw,
                                                              int Hin, int Win, int B, const float* __restrict__ aff) {
  constexpr int V = Vec<T>::N;
  constexpr int CH = FirDownCh<T>::value;
  constexpr int VPC = CH / V;  // 4
  constexpr int TO = 8, WIN = 18, NPX = WIN * WIN;
  // Shared-memory layout: a pixel's slice is 64 B = half of the 32 banks, and the filter reads every OTHER pixel of a
  // row (8 of them per warp-wide LDS.128): with a plain [row][pixel] layout all eight land in the same half (2x bank
  // conflicts, and shared-memory wavefronts were this kernel's limiter).  Pixels 2 and 3 of every group of four are
  // swapped (physical column = c ^ ((c >> 1) & 1)), which makes the bank half alternate along any stride-2 walk; rows
  // are padded to 20 pixels (a multiple of 128 B).
  constexpr int WROW = 20, NPHYS = WIN * WROW;
  extern __shared__ __align__(16) unsigned char smraw[];
  T* sa = reinterpret_cast<T*>(smraw);  // activated [WIN][WROW][CH]
  T* sr = sa + NPHYS * CH;              // raw       [WIN][WROW][CH]
  float* saff = reinterpret_cast<float*>(sr + NPHYS * CH);
  auto phys = [](int wr, int wc) { return wr * WROW + (wc ^ ((wc >> 1) & 1)); };
  const int C = s0.C;
  const int Hout = Hin / 2, Wout = Win / 2;
  const int tiles_x = (Wout + TO - 1) / TO;
  const int v = threadIdx.x & (VPC - 1), ps = threadIdx.x / VPC;
  constexpr int PSTEP = 256 / VPC;                 // 64
  constexpr int NIT = (NPX + PSTEP - 1) / PSTEP;   // 6
  // One tile per block; grid = (channel slice, tile, sample): the slices of the same pixels run concurrently, so DRAM
  // sees whole rows.  (Measured alternatives at 512 x 640 x 128 x 16 clips: persistent blocks with a register-level
  // prefetch of the next window 1.7 TB/s, 128-byte slices with 2x2 outputs per thread 1.8 TB/s, this form 2.0 TB/s and,
  // with the conflict-free shared-memory layout below, 2.7 TB/s in fp32.)
  {
    const int b = blockIdx.z, c0 = blockIdx.x * CH;
    const int ty = blockIdx.y / tiles_x;
    const int oy0 = ty * TO, ox0 = (blockIdx.y - ty * tiles_x) * TO;
    const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
    const T* src = s0.x + static_cast<size_t>(b) * Hin * Win * C + c0 + v * V;
    // interior tiles (all but the image border) need no bounds checks and no zero padding
    const bool interior = iy0 >= 0 && ix0 >= 0 && iy0 + WIN <= Hin && ix0 + WIN <= Win;
    uint4 cur[NIT];
    uint32_t inb_cur = 0;
    if (interior) {
      const T* s00 = src + (static_cast<size_t>(iy0) * Win + ix0) * C;
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int pw = ps + k * PSTEP;
        const int wr = pw / WIN, wc = pw - wr * WIN;
        const bool in = pw < NPX;
        inb_cur |= in ? (1u << k) : 0u;
        cur[k] = in ? __ldg(reinterpret_cast<const uint4*>(s00 + (static_cast<size_t>(wr) * Win + wc) * C)) : make_uint4(0, 0, 0, 0);
      }
    } else {
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int pw = ps + k * PSTEP;
        const int wr = pw / WIN, wc = pw - wr * WIN;
        const int iy = iy0 + wr, ix = ix0 + wc;
        const bool in = pw < NPX && iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
        inb_cur |= in ? (1u << k) : 0u;
        cur[k] = in ? __ldg(reinterpret_cast<const uint4*>(src + (static_cast<size_t>(iy) * Win + ix) * C)) : make_uint4(0, 0, 0, 0);
      }
    }
    if (threadIdx.x < CH) {
      const int c = c0 + threadIdx.x;
      if (aff != nullptr) {
        // scale / shift precomputed once per launch (launch_gn_affine)
        saff[threadIdx.x] = __ldg(aff + (static_cast<size_t>(b) * 2) * C + c);
        saff[CH + threadIdx.x] = __ldg(aff + (static_cast<size_t>(b) * 2 + 1) * C + c);
      } else {
        const int G = min(C / 4, 32), cpg = C / G;
        const int g = c / cpg;
        const double inv_cnt = 1.0 / (static_cast<double>(Hin) * Win * cpg);
        double sum = 0.0, sq = 0.0;
        for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {
          const longlong2 st = __ldg(reinterpret_cast<const longlong2*>(s0.stats + (static_cast<size_t>(b) * C + cc) * 2));
          sum += static_cast<double>(st.x) * (1.0 / kStatSumScale);
          sq += static_cast<double>(st.y) * (1.0 / kStatSqScale);
        }
        const double mean = sum * inv_cnt;
        double var = sq * inv_cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = rsqrtf(static_cast<float>(var) + eps);
        const float sc = gamma[c] * rstd;
        saff[threadIdx.x] = sc;
        saff[CH + threadIdx.x] = beta[c] - static_cast<float>(mean) * sc;
      }
    }
    __syncthreads();  // saff ready; the previous tile's filter phase is done with sa / sr
    // ---- phase 1: activate the window (each input element once) ----
    {
      float sc[V], sh[V];
#pragma unroll
      for (int j = 0; j < V; ++j) { sc[j] = saff[v * V + j]; sh[j] = saff[CH + v * V + j]; }
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int pw = ps + k * PSTEP;
        if (pw < NPX) {
          float f[V], a[V];
          Vec<T>::unpack(cur[k], f);
          const bool in = (inb_cur >> k) & 1u;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const float n = fmaf(f[j], sc[j], sh[j]);
            a[j] = in ? (do_silu ? silu_act<T>(n) : n) : 0.f;  // the FIR pads the ACTIVATED tensor with zeros
          }
          const int wr = pw / WIN, wc = pw - wr * WIN;
          const int pp = phys(wr, wc);
          sts_vec<T>(sa + pp * CH + v * V, a);
          *reinterpret_cast<uint4*>(sr + pp * CH + v * V) = cur[k];
        }
      }
    }
    __syncthreads();
    // ---- phase 2: item = (output pixel, 4-channel quad); out[i][j] = sum_{a,b<4} k[a] k[b] win[2i+a][2j+b] ----
    constexpr int NQ = CH / 4;
    for (int it = threadIdx.x; it < TO * TO * NQ; it += 256) {
      const int q = it % NQ, px = it / NQ;
      const int oy = px / TO, ox = px - oy * TO;
      // window column 2 ox + bb -> physical column c ^ ((c >> 1) & 1)
      const T* pa = sa + ((2 * oy) * WROW) * CH + q * 4;
      const T* pr = sr + ((2 * oy) * WROW) * CH + q * 4;
      int colp[4];
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) { const int c = 2 * ox + bb; colp[bb] = (c ^ ((c >> 1) & 1)) * CH; }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), raw = acc;
      const float k1[4] = {0.125f, 0.375f, 0.375f, 0.125f};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float4 ha = make_float4(0.f, 0.f, 0.f, 0.f), hr = ha;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          fma4(ha, k1[bb], lds4<T>(pa + a * WROW * CH + colp[bb]));
          fma4(hr, k1[bb], lds4<T>(pr + a * WROW * CH + colp[bb]));
        }
        fma4(acc, k1[a], ha);
        fma4(raw, k1[a], hr);
      }
      const int gy = oy0 + oy, gx = ox0 + ox;
      if (gy < Hout && gx < Wout) {
        const size_t o = ((static_cast<size_t>(b) * Hout + gy) * Wout + gx) * C + c0 + q * 4;
        auto put = [&](T* dst, const float4& v4) {
          if constexpr (DT<T>::kIsBf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v4.x, v4.y), hi = __floats2bfloat162_rn(v4.z, v4.w);
            *reinterpret_cast<uint2*>(dst + o) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
          } else {
            *reinterpret_cast<float4*>(dst + o) =
                as_operand ? make_float4(round_tf32(v4.x), round_tf32(v4.y), round_tf32(v4.z), round_tf32(v4.w)) : v4;
          }
        };
        put(out_act, acc);
        if (out_raw != nullptr) put(out_raw, raw);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// normalise + SiLU + FIR x2 DOWN (single source), third generation (round 2): a STREAMING form without shared memory.
// The tiled kernel above is issue-bound (ncu: ~840 instructions per thread per 8 x 8 tile, 39 % of the HBM roofline):
// every input element costs a global load, an activation, a shared-memory store and four shared-memory loads, plus
// two block-wide barriers per tile and a 27 % halo.  Here a WARP walks down a strip of 16 input columns x 64 bytes of
// channels: lane = (column pair pp = lane >> 2, 16-byte channel vector q = lane & 3); a lane owns input columns
// (2j - 1, 2j) of output column j = j0 + pp, activates them ONCE, gets columns (2j + 1, 2j + 2) from lane + 4 by
// shuffle (lane pp = 7 has no right neighbour: 7 outputs per 16 columns, 14 % column halo, no row halo inside a
// segment), filters horizontally in registers, and keeps TWO running vertical accumulators (output rows i - 1 and i)
// instead of a window: input row 2i - 1 adds k2 h to output i - 1 and starts output i with k0 h, row 2i adds k3 h /
// k1 h and emits output i - 1.  No shared memory, no barriers, ~270 instructions per output pixel x 4 channels (was
// ~840), same FMA order as the tiled kernel (bit-identical results; in bf16 mode the activated value is rounded to
// bf16 before the filter, as the tiled kernel's staging did).
// ------------------------------------------------------------------------------------------------
constexpr int kFirSegRows = 32;  // output rows per warp task (halved down to 8 while the launch has < ~3 waves of warps)
constexpr int kFirDepth = 4;     // loop iterations (pairs of input rows) in flight per warp
constexpr int kFirSmem = 4 * kFirDepth * 4 * 32 * 16;  // 4 warps x depth x 4 pixels x 32 lanes x 16 B = 32 KB

template <typename T>
__global__ void __launch_bounds__(128, DT<T>::kIsBf16 ? 4 : 5) gn_fir_down_stream_kernel(GnSrcT<T> s0, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float eps, int do_silu,
                                                                  int as_operand, T* __restrict__ out_act,
                                                                  T* __restrict__ out_raw, int Hin, int Win, int B,
                                                                  const float* __restrict__ aff, int nstrips, int nsegs,
                                                                  int seg_rows, long long ntasks) {
  constexpr int V = Vec<T>::N;
  constexpr int CH = 4 * V;  // channels per warp (64 bytes)
  extern __shared__ __align__(16) uint4 ring[];
  const int lane = threadIdx.x & 31;
  const int q = lane & 3, pp = lane >> 2;
  const int C = s0.C;
  const int nsl = C / CH;
  const int Hout = Hin >> 1, Wout = Win >> 1;
  const float k1[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (long long task = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5); task < ntasks;
       task += static_cast<long long>(gridDim.x) * 4) {
    // channel slice fastest: the slices of the same pixels run side by side, so DRAM sees whole pixel rows
    long long t = task;
    const int sl = static_cast<int>(t % nsl); t /= nsl;
    const int strip = static_cast<int>(t % nstrips); t /= nstrips;
    const int seg = static_cast<int>(t % nsegs);
    const int b = static_cast<int>(t / nsegs);
    const int c0 = sl * CH + q * V;
    const int j = strip * 7 + pp;              // output column of this lane (lane pp = 7: halo only)
    const int xa = 2 * j - 1, xb = 2 * j;      // owned input columns
    const bool col_a = xa >= 0 && xa < Win, col_b = xb < Win;
    const bool emit = pp < 7 && j < Wout;
    const int i0 = seg * seg_rows, i1 = min(i0 + seg_rows, Hout);
    // per-channel scale / shift of this lane's V channels
    float sc[V], sh[V];
    if (aff != nullptr) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        sc[v] = __ldg(aff + (static_cast<size_t>(b) * 2) * C + c0 + v);
        sh[v] = __ldg(aff + (static_cast<size_t>(b) * 2 + 1) * C + c0 + v);
      }
    } else {
      const int G = min(C / 4, 32), cpg = C / G;
      const double inv_cnt = 1.0 / (static_cast<double>(Hin) * Win * cpg);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int c = c0 + v, g = c / cpg;
        double sum = 0.0, sq = 0.0;
        for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {
          const longlong2 st = __ldg(reinterpret_cast<const longlong2*>(s0.stats + (static_cast<size_t>(b) * C + cc) * 2));
          sum += static_cast<double>(st.x) * (1.0 / kStatSumScale);
          sq += static_cast<double>(st.y) * (1.0 / kStatSqScale);
        }
        const double mean = sum * inv_cnt;
        double var = sq * inv_cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = rsqrtf(static_cast<float>(var) + eps);
        sc[v] = gamma[c] * rstd;
        sh[v] = beta[c] - static_cast<float>(mean) * sc[v];
      }
    }
    const T* src = s0.x + static_cast<size_t>(b) * Hin * Win * C + c0;
    float acc_a[V], acc_r[V];  // running output row (activated / raw)
#pragma unroll
    for (int v = 0; v < V; ++v) acc_a[v] = acc_r[v] = 0.f;

    // Prefetch ring: the owned pixels of the rows of the next kFirDepth iterations travel global -> shared memory by
    // cp.async (zero-filled outside the image), each lane into its OWN 16-byte column of the ring (conflict-free, and no
    // cross-lane hand-off: a lane reads back only what it requested).  ncu on the register-prefetch form: 47 % of the
    // stall cycles on global loads at 12-20 resident warps per SM; the ring hides them without spending registers.
    auto issue_row = [&](int r, uint4* dst_a, uint4* dst_b) {
      const bool rin = r >= 0 && r < Hin;
      const T* rowp = src + static_cast<size_t>(rin ? r : 0) * Win * C;
      const T* pa_ = rowp + static_cast<size_t>(col_a ? xa : 0) * C;
      const T* pb_ = rowp + static_cast<size_t>(col_b ? xb : 0) * C;
      const unsigned na = (rin && col_a) ? 16u : 0u, nb = (rin && col_b) ? 16u : 0u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_a)), "l"(pa_), "r"(na) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_b)), "l"(pb_), "r"(nb) : "memory");
    };
    // horizontally filtered row r -> ha (activated), hr (raw); zero outside the image (the FIR pads the ACTIVATED tensor)
    auto hrow = [&](int r, const uint4& ra, const uint4& rb, float (&ha)[V], float (&hr)[V]) {
      const bool rin = r >= 0 && r < Hin;
      float fa[V], fb[V], aa[V], ab[V];
      Vec<T>::unpack(ra, fa);
      Vec<T>::unpack(rb, fb);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float na = fmaf(fa[v], sc[v], sh[v]), nb = fmaf(fb[v], sc[v], sh[v]);
        aa[v] = (rin && col_a) ? (do_silu ? silu_act<T>(na) : na) : 0.f;
        ab[v] = (rin && col_b) ? (do_silu ? silu_act<T>(nb) : nb) : 0.f;
      }
      // the activated values pass through the act dtype (bf16 mode: rounded, exactly like a materialised tensor)
      uint4 pa, pb;
      if constexpr (DT<T>::kIsBf16) {
        pa = Vec<T>::pack_operand(aa);
        pb = Vec<T>::pack_operand(ab);
        Vec<T>::unpack(pa, aa);
        Vec<T>::unpack(pb, ab);
      } else {
        pa = make_uint4(__float_as_uint(aa[0]), __float_as_uint(aa[1]), __float_as_uint(aa[2]), __float_as_uint(aa[3]));
        pb = make_uint4(__float_as_uint(ab[0]), __float_as_uint(ab[1]), __float_as_uint(ab[2]), __float_as_uint(ab[3]));
      }
      // columns 2j + 1, 2j + 2 = the owned columns of lane + 4
      auto shfl4 = [](const uint4& x) {
        return make_uint4(__shfl_down_sync(0xffffffffu, x.x, 4), __shfl_down_sync(0xffffffffu, x.y, 4),
                          __shfl_down_sync(0xffffffffu, x.z, 4), __shfl_down_sync(0xffffffffu, x.w, 4));
      };
      const uint4 qa = shfl4(pa), qb = shfl4(pb), sa = shfl4(ra), sb = shfl4(rb);
      float ac[V], ad[V], fc[V], fd[V];
      Vec<T>::unpack(qa, ac);
      Vec<T>::unpack(qb, ad);
      Vec<T>::unpack(sa, fc);
      Vec<T>::unpack(sb, fd);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        ha[v] = fmaf(k1[3], ad[v], fmaf(k1[2], ac[v], fmaf(k1[1], ab[v], k1[0] * aa[v])));
        hr[v] = fmaf(k1[3], fd[v], fmaf(k1[2], fc[v], fmaf(k1[1], fb[v], k1[0] * fa[v])));
      }
    };

    uint4* my = ring + static_cast<size_t>(threadIdx.x >> 5) * (kFirDepth * 4 * 32) + lane;  // entry (slot, k): my[(slot * 4 + k) * 32]
#pragma unroll
    for (int k = 0; k < kFirDepth; ++k) {
      if (i0 + k <= i1) {
        issue_row(2 * (i0 + k) - 1, my + (k * 4 + 0) * 32, my + (k * 4 + 1) * 32);
        issue_row(2 * (i0 + k), my + (k * 4 + 2) * 32, my + (k * 4 + 3) * 32);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int i = i0; i <= i1; ++i) {
      const int slot = (i - i0) % kFirDepth;
      asm volatile("cp.async.wait_group %0;" ::"n"(kFirDepth - 1) : "memory");
      const uint4 c0a = my[(slot * 4 + 0) * 32], c0b = my[(slot * 4 + 1) * 32], c1a = my[(slot * 4 + 2) * 32],
                  c1b = my[(slot * 4 + 3) * 32];
      if (i + kFirDepth <= i1) {
        issue_row(2 * (i + kFirDepth) - 1, my + (slot * 4 + 0) * 32, my + (slot * 4 + 1) * 32);
        issue_row(2 * (i + kFirDepth), my + (slot * 4 + 2) * 32, my + (slot * 4 + 3) * 32);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      float ha0[V], hr0[V], ha1[V], hr1[V];
      hrow(2 * i - 1, c0a, c0b, ha0, hr0);
      hrow(2 * i, c1a, c1b, ha1, hr1);
      if (i > i0) {
        // finish output row i - 1: taps 2 and 3
        float oa[V], orr[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          oa[v] = fmaf(k1[3], ha1[v], fmaf(k1[2], ha0[v], acc_a[v]));
          orr[v] = fmaf(k1[3], hr1[v], fmaf(k1[2], hr0[v], acc_r[v]));
        }
        if (emit) {
          const size_t o = ((static_cast<size_t>(b) * Hout + (i - 1)) * Wout + j) * C + c0;
          if (as_operand) Vec<T>::store_operand(out_act + o, oa);
          else Vec<T>::store(out_act + o, oa);
          if (out_raw != nullptr) {
            if (as_operand) Vec<T>::store_operand(out_raw + o, orr);
            else Vec<T>::store(out_raw + o, orr);
          }
        }
      }
      // start output row i: taps 0 and 1
#pragma unroll
      for (int v = 0; v < V; ++v) {
        acc_a[v] = fmaf(k1[1], ha1[v], k1[0] * ha0[v]);
        acc_r[v] = fmaf(k1[1], hr1[v], k1[0] * hr0[v]);
      }
    }
  }
}

void launch_gn_apply(int dt, GnSrc s0, GnSrc s1, const float* gamma, const float* beta, float eps, int fir, bool do_silu,
                     bool as_operand, void* out_act, void* out_raw, int B, int Hin, int Win, cudaStream_t st,
                     const float* aff) {
  const int Ct = s0.C + s1.C;
  const int Hout = fir == 1 ? Hin / 2 : (fir == 2 ? Hin * 2 : Hin);
  const int Wout = fir == 1 ? Win / 2 : (fir == 2 ? Win * 2 : Win);
  if (fir != 0 && s1.C == 0 && s0.C % (dt == kBF16 ? 64 : 32) == 0) {
    DISPATCH_DT(dt, {
      constexpr int CH = FirCh<T>::value;
      GnSrcT<T> a{(const T*)s0.x, s0.stats, s0.C};
      static const bool tiled_down = getenv("USE_B200_FIR_DOWN") && getenv("USE_B200_FIR_DOWN")[0] == 't';  // A/B switch
      if (fir == 1 && !tiled_down) {
        const int nstrips = (Wout + 6) / 7;
        int seg_rows = kFirSegRows;
        auto tasks_for = [&](int rows) { return static_cast<long long>(s0.C / (4 * Vec<T>::N)) * nstrips * ((Hout + rows - 1) / rows) * B; };
        while (seg_rows > 8 && tasks_for(seg_rows) < 148LL * 20 * 3) seg_rows >>= 1;
        const int nsegs = (Hout + seg_rows - 1) / seg_rows;
        const long long ntasks = tasks_for(seg_rows);
        const int blocks = static_cast<int>(std::min<long long>((ntasks + 3) / 4, 148LL * 128));
        gn_fir_down_stream_kernel<T><<<blocks, 128, kFirSmem, st>>>(a, gamma, beta, eps, do_silu, as_operand, (T*)out_act,
                                                             (T*)out_raw, Hin, Win, B, aff, nstrips, nsegs, seg_rows, ntasks);
      } else if (fir == 1) {
        constexpr int CHD = FirDownCh<T>::value;
        dim3 grid(s0.C / CHD, ((Hout + 7) / 8) * ((Wout + 7) / 8), B);
        const size_t sm = 2 * 18 * 20 * CHD * sizeof(T) + 2 * CHD * sizeof(float);
        gn_fir_down_kernel<T><<<grid, 256, sm, st>>>(a, gamma, beta, eps, do_silu, as_operand, (T*)out_act, (T*)out_raw, Hin, Win, B, aff);
      } else {
        dim3 grid(((Hout + 15) / 16) * ((Wout + 15) / 16), s0.C / CH, B);
        const size_t sm = 2 * 10 * 10 * CH * sizeof(T) + 2 * CH * sizeof(float);
        gn_apply_fir_tiled_kernel<T, 2><<<grid, 256, sm, st>>>(a, gamma, beta, eps, do_silu, as_operand, (T*)out_act,
                                                              (T*)out_raw, Hin, Win, aff);
      }
    });
    return;
  }
  DISPATCH_DT(dt, {
    constexpr int V = Vec<T>::N;
    const int vpp = Ct / V;
    const int threads = vpp >= 256 ? vpp : (256 / vpp) * vpp;  // a multiple of vpp (<= 1024 for Ct <= 4096)
    const int nwork = fir == 0 ? Hout * Wout : (fir == 2 ? Hin * Win : ((Hout + 1) / 2) * ((Wout + 1) / 2));
    // enough blocks for >= ~8 waves of 148 SMs when the tensor allows it, at most kGnMaxWorkPerBlock items per block
    const int rows = threads / vpp;
    long long wpb = (static_cast<long long>(nwork) * B + 148 * 8 - 1) / (148 * 8);
    wpb = std::max<long long>(rows, std::min<long long>(kGnMaxWorkPerBlock, (wpb + rows - 1) / rows * rows));
    const int work_per_block = static_cast<int>(wpb);
    dim3 grid((nwork + work_per_block - 1) / work_per_block, B);
    GnSrcT<T> a{(const T*)s0.x, s0.stats, s0.C};
    GnSrcT<T> c{(const T*)s1.x, s1.stats, s1.C};
    const size_t sm = 2 * Ct * sizeof(double) + 2 * Ct * sizeof(float);
    if (fir == 0)
      gn_apply_kernel<T, 0><<<grid, threads, sm, st>>>(a, c, gamma, beta, eps, do_silu, as_operand, (T*)out_act, (T*)out_raw, Hin, Win, work_per_block);
    else if (fir == 1)
      gn_apply_kernel<T, 1><<<grid, threads, sm, st>>>(a, c, gamma, beta, eps, do_silu, as_operand, (T*)out_act, (T*)out_raw, Hin, Win, work_per_block);
    else
      gn_apply_kernel<T, 2><<<grid, threads, sm, st>>>(a, c, gamma, beta, eps, do_silu, as_operand, (T*)out_act, (T*)out_raw, Hin, Win, work_per_block);
  });
}

// ------------------------------------------------------------------------------------------------
// input convolution: 3x3 pad 1, 4 -> N channels.  fp32 NHWC4 input, act-dtype output.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv_in4_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, T* __restrict__ out, int H, int W,
                                                        int N, long long total) {
  constexpr int V = Vec<T>::N;
  extern __shared__ float sw[];  // [36][N] (tap*4+ci major), then bias [N]
  for (int i = threadIdx.x; i < 36 * N; i += blockDim.x) {
    const int co = i % N, k = i / N;  // k = tap*4 + ci
    const int tap = k >> 2, ci = k & 3;
    sw[i] = w[(co * 4 + ci) * 9 + tap];
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) sw[36 * N + i] = bias[i];
  __syncthreads();
  const int vpp = N / V;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % vpp);
    const long long pix = i / vpp;  // over B*H*W
    const int xw = static_cast<int>(pix % W);
    const int yh = static_cast<int>((pix / W) % H);
    const long long b = pix / (static_cast<long long>(W) * H);
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = sw[36 * N + cv * V + j];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = yh + r - 1;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = xw + s - 1;
        if (ix < 0 || ix >= W) continue;
        const float4 in = __ldg(reinterpret_cast<const float4*>(x) + (b * H + iy) * W + ix);
        const float iv[4] = {in.x, in.y, in.z, in.w};
        const int tap = r * 3 + s;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const float* wr = sw + (tap * 4 + ci) * N + cv * V;
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] += iv[ci] * wr[j];
        }
      }
    }
    Vec<T>::store(out + pix * N + cv * V, acc);
  }
}

void launch_conv_in4(int dt, const float* x, const float* w, const float* bias, void* out, int B, int H, int W, int N,
                     cudaStream_t st) {
  DISPATCH_DT(dt, {
    constexpr int V = Vec<T>::N;
    const long long total = static_cast<long long>(B) * H * W * (N / V);
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 32));
    conv_in4_kernel<T><<<blocks, 256, (37 * N) * sizeof(float), st>>>(x, w, bias, (T*)out, H, W, N, total);
  });
}

// ------------------------------------------------------------------------------------------------
// pyramid convolution: 3x3 pad 1, C -> 4 channels (fp32 out), optional + FIR-upsample(prev pyramid).
// One thread = PX horizontally adjacent pixels x 4 outputs; weights [tap][c][4] in shared memory.
// ------------------------------------------------------------------------------------------------
template <typename T, int PX>
__global__ void __launch_bounds__(128) conv_out4_kernel(const T* __restrict__ a, const float* __restrict__ w,
                                                         const float* __restrict__ bias, const float* __restrict__ prev,
                                                         float* __restrict__ out, int H, int W, int C, long long total) {
  constexpr int V = Vec<T>::N;
  extern __shared__ float4 sw4[];  // [9][C] float4 over the 4 output channels
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
    const int c = i % C, tap = i / C;
    sw4[i] = make_float4(w[(0 * C + c) * 9 + tap], w[(1 * C + c) * 9 + tap], w[(2 * C + c) * 9 + tap],
                         w[(3 * C + c) * 9 + tap]);
  }
  __syncthreads();
  const int wq = (W + PX - 1) / PX;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xq = static_cast<int>(i % wq) * PX;
    const int yh = static_cast<int>((i / wq) % H);
    const long long b = i / (static_cast<long long>(wq) * H);
    float acc[PX][4];
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      acc[p][0] = bias[0]; acc[p][1] = bias[1]; acc[p][2] = bias[2]; acc[p][3] = bias[3];
    }
    for (int r = 0; r < 3; ++r) {
      const int iy = yh + r - 1;
      if (iy < 0 || iy >= H) continue;
      const T* row = a + ((b * H + iy) * W) * C;
      for (int c = 0; c < C; c += V) {
        // input columns xq-1 .. xq+PX
        float f[PX + 2][V];
#pragma unroll
        for (int q = 0; q < PX + 2; ++q) {
          const int ix = xq + q - 1;
          if (ix >= 0 && ix < W) {
            Vec<T>::load(row + static_cast<size_t>(ix) * C + c, f[q]);
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) f[q][j] = 0.f;
          }
        }
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const float4* wr = sw4 + (r * 3 + s) * C + c;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const float4 ww = wr[j];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
              const float v = f[p + s][j];
              acc[p][0] += v * ww.x; acc[p][1] += v * ww.y; acc[p][2] += v * ww.z; acc[p][3] += v * ww.w;
            }
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      const int ox = xq + p;
      if (ox >= W) continue;
      if (prev != nullptr) {
        const int Hp = H / 2, Wp = W / 2;
        const int my = yh >> 1, mx = ox >> 1;
        const int ya = (yh & 1) ? my : my - 1, xa = (ox & 1) ? mx : mx - 1;
        const float wya = (yh & 1) ? 0.75f : 0.25f, wxa = (ox & 1) ? 0.75f : 0.25f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const int yy = ya + dy;
          if (yy < 0 || yy >= Hp) continue;
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int xx = xa + dx;
            if (xx < 0 || xx >= Wp) continue;
            const float kw = (dy ? 1.f - wya : wya) * (dx ? 1.f - wxa : wxa);
            const float4 pv = __ldg(reinterpret_cast<const float4*>(prev) + (b * Hp + yy) * Wp + xx);
            acc[p][0] += kw * pv.x; acc[p][1] += kw * pv.y; acc[p][2] += kw * pv.z; acc[p][3] += kw * pv.w;
          }
        }
      }
      reinterpret_cast<float4*>(out)[(b * H + yh) * W + ox] = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
    }
  }
}

void launch_conv_out4(int dt, const void* a, const float* w, const float* bias, const float* prev, float* out, int B,
                      int H, int W, int C, cudaStream_t st) {
  constexpr int PX = 4;
  const long long total = static_cast<long long>(B) * H * ((W + PX - 1) / PX);
  const int blocks = static_cast<int>(std::min<long long>((total + 127) / 128, 148LL * 16));
  DISPATCH_DT(dt, {
    auto kern = conv_out4_kernel<T, PX>;
    const size_t sm = static_cast<size_t>(9) * C * sizeof(float4);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
    kern<<<blocks, 128, sm, st>>>((const T*)a, w, bias, prev, out, H, W, C, total);
  });
}

// ------------------------------------------------------------------------------------------------
// Combine(sum): out = h + bias + W[C][pc] . pyr  (layerspp.py:50-55), optionally with the per-channel GroupNorm
// statistics of `out` (fixed point, like the conv epilogue) so the next ResBlock needs no separate statistics pass.
// One thread owns a 16-byte channel vector (its weights and bias live in registers) and walks over the pixels of the
// block's fixed range: the block partition does not depend on the batch, so the statistics are batch-invariant.
// ------------------------------------------------------------------------------------------------
constexpr int kCombinePixPerBlock = 256;

// four consecutive channels of T <-> floats (bf16: one 8-byte access).  The Combine kernel uses 4-channel vectors for BOTH
// dtypes: with 8 (the 16-byte bf16 vector) it needed 128 registers and ran at 18 % of the HBM roofline (ncu r02: 24 %
// issue-active, two blocks per SM) against 68 % for the fp32 instantiation.
template <typename T> struct Quad;
template <> struct Quad<float> {
  __device__ __forceinline__ static void load(const float* p, float (&v)[4]) { Vec<float>::load(p, v); }
  __device__ __forceinline__ static void store(float* p, const float (&v)[4]) { Vec<float>::store(p, v); }
};
template <> struct Quad<__nv_bfloat16> {
  __device__ __forceinline__ static void load(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  }
  __device__ __forceinline__ static void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
};

template <typename T, int PC>
__global__ void __launch_bounds__(256) combine_kernel(const T* __restrict__ h, const float* __restrict__ pyr,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       T* __restrict__ out, long long* __restrict__ stats, int HW, int C,
                                                       int pix_per_block) {
  constexpr int V = 4;
  extern __shared__ float sred[];  // [rows][C][2]
  const int vpp = C / V;
  const int rows = blockDim.x / vpp;
  const int cv = threadIdx.x % vpp, prow = threadIdx.x / vpp;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  float wv[V][PC], bv[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    bv[j] = bias[cv * V + j];
#pragma unroll
    for (int k = 0; k < PC; ++k) wv[j][k] = w[(cv * V + j) * PC + k];
  }
  float s[V], q[V];
#pragma unroll
  for (int j = 0; j < V; ++j) s[j] = q[j] = 0.f;
  const size_t base = static_cast<size_t>(b) * HW;
  if (prow < rows) {
    constexpr int U = 4;  // independent pixel loads in flight per thread (a serial loop here is pure load latency; 8 costs
                          // 128 registers and a block per SM: no gain)
    for (int pb = p0 + prow; pb < p1; pb += U * rows) {
      float pv[U][PC], f[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = pb + u * rows;
        if (p < p1) {
          if constexpr (PC == 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(pyr) + base + p);
            pv[u][0] = t.x; pv[u][1] = t.y; pv[u][2] = t.z; pv[u][3] = t.w;
          } else {
            const float2 t = __ldg(reinterpret_cast<const float2*>(pyr) + base + p);
            pv[u][0] = t.x; pv[u][1] = t.y;
          }
          Quad<T>::load(h + (base + p) * C + cv * V, f[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = pb + u * rows;
        if (p < p1) {
#pragma unroll
          for (int j = 0; j < V; ++j) {
            float acc = bv[j];
            if constexpr (PC == 4) acc += wv[j][0] * pv[u][0] + wv[j][1] * pv[u][1] + wv[j][2] * pv[u][2] + wv[j][3] * pv[u][3];
            else acc += wv[j][0] * pv[u][0] + wv[j][1] * pv[u][1];
            f[u][j] = acc + f[u][j];  // same association as conv2d then "+ h": (bias + w.p) + h
            s[j] += f[u][j];
            q[j] = fmaf(f[u][j], f[u][j], q[j]);
          }
          Quad<T>::store(out + (base + p) * C + cv * V, f[u]);
        }
      }
    }
  }
  if (stats == nullptr) return;
  if (prow < rows) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      sred[(prow * C + cv * V + j) * 2] = s[j];
      sred[(prow * C + cv * V + j) * 2 + 1] = q[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double a = 0.0, qq = 0.0;
    for (int r = 0; r < rows; ++r) {  // fixed order: deterministic
      a += static_cast<double>(sred[(r * C + c) * 2]);
      qq += static_cast<double>(sred[(r * C + c) * 2 + 1]);
    }
    stat_atomic_add(stats + (static_cast<size_t>(b) * C + c) * 2, a, qq);
  }
}

void launch_combine(int dt, const void* h, const float* pyr, const float* w, const float* bias, void* out, long long* stats,
                    int B, int HW, int C, int pc, cudaStream_t st) {
  DISPATCH_DT(dt, {
    constexpr int V = 4;
    const int vpp = C / V;
    const int threads = vpp >= 256 ? vpp : (256 / vpp) * vpp;
    const int rows = threads / vpp;
    // pixels per block: never a function of the batch, so the per-block partial sums -- and with them the fixed-point
    // statistics -- do not depend on how many clips are in flight.  (1024-pixel blocks were measured: no gain in fp32,
    // slower in bf16 -- 320 blocks are barely one wave.)
    // The low-resolution levels (<= 64 x 80) take shorter blocks: with 256 pixels a thread walks 32-64 pixels serially and
    // a 2-block launch is pure load latency (batch 1: 34 us for 0.3 MB); a function of the LEVEL only, never of the batch.
    const int ppb = HW >= 16384 ? kCombinePixPerBlock : (HW >= 4096 ? 64 : 32);
    dim3 grid((HW + ppb - 1) / ppb, B);
    const size_t sm = static_cast<size_t>(rows) * C * 2 * sizeof(float);
    if (pc == 4) combine_kernel<T, 4><<<grid, threads, sm, st>>>((const T*)h, pyr, w, bias, (T*)out, stats, HW, C, ppb);
    else combine_kernel<T, 2><<<grid, threads, sm, st>>>((const T*)h, pyr, w, bias, (T*)out, stats, HW, C, ppb);
  });
}

// ------------------------------------------------------------------------------------------------
// FIR downsample of the pc-channel fp32 input pyramid (one thread = one output pixel channel)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fir4_down_kernel(const float* __restrict__ x, float* __restrict__ out, int Hin,
                                                         int Win, int pc, long long total) {
  const int Ho = Hin / 2, Wo = Win / 2;
  const float k1[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % pc);
    const long long pix = i / pc;
    const int ox = static_cast<int>(pix % Wo), oy = static_cast<int>((pix / Wo) % Ho);
    const long long b = pix / (static_cast<long long>(Wo) * Ho);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int iy = 2 * oy + a - 1;
      if (iy < 0 || iy >= Hin) continue;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int ix = 2 * ox + bb - 1;
        if (ix < 0 || ix >= Win) continue;
        acc += (k1[a] * k1[bb]) * __ldg(x + ((b * Hin + iy) * Win + ix) * pc + c);
      }
    }
    out[i] = acc;
  }
}

void launch_fir4_down(const float* x, float* out, int B, int Hin, int Win, int pc, cudaStream_t st) {
  const long long total = static_cast<long long>(B) * (Hin / 2) * (Win / 2) * pc;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 32));
  fir4_down_kernel<<<blocks, 256, 0, st>>>(x, out, Hin, Win, pc, total);
}

// ------------------------------------------------------------------------------------------------
// generic direct convolution: verification of the tcgen05 path only (slow by construction)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void conv_ref_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                int bias_bstride, const T* __restrict__ res, float scale, T* __restrict__ out, int H, int W,
                                int Cin, int Cout, int ks, long long total) {
  const int pad = ks / 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % Cout);
    const long long pix = i / Cout;
    const int xw = static_cast<int>(pix % W), yh = static_cast<int>((pix / W) % H);
    const long long b = pix / (static_cast<long long>(W) * H);
    float acc = 0.f;
    for (int r = 0; r < ks; ++r) {
      const int iy = yh + r - pad;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < ks; ++s) {
        const int ix = xw + s - pad;
        if (ix < 0 || ix >= W) continue;
        const T* px = x + ((b * H + iy) * W + ix) * Cin;
        const float* pw = w + (static_cast<size_t>(co) * Cin) * ks * ks + r * ks + s;
        for (int ci = 0; ci < Cin; ++ci) acc += static_cast<float>(px[ci]) * pw[static_cast<size_t>(ci) * ks * ks];
      }
    }
    acc += bias[b * bias_bstride + co];
    if (res != nullptr) acc += static_cast<float>(res[i]);
    out[i] = static_cast<T>(acc * scale);
  }
}

void launch_conv_ref(int dt, const void* x, const float* w, const float* bias, int bias_bstride, const void* res,
                     float scale, void* out, int B, int H, int W, int Cin, int Cout, int ksize, cudaStream_t st) {
  const long long total = static_cast<long long>(B) * H * W * Cout;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 64));
  DISPATCH_DT(dt, {
    conv_ref_kernel<T><<<blocks, 256, 0, st>>>((const T*)x, w, bias, bias_bstride, (const T*)res, scale, (T*)out, H, W, Cin,
                                               Cout, ksize, total);
  });
}

// ------------------------------------------------------------------------------------------------
// 3xTF32 parity mode (USE_DTYPE_F32X3): split an fp32 channel window into two TF32 operands,
// hi = rne_tf32(x), lo = rne_tf32(x - hi)  (x - hi is exact in fp32), so that
// x w ~= x_hi w_hi + x_hi w_lo + x_lo w_hi with a relative error of ~2^-21 instead of TF32's 2^-11.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, int Ct, int c0, int C,
                                                          float* __restrict__ hi, float* __restrict__ lo, size_t npix) {
  const size_t n = npix * static_cast<size_t>(C);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = i / C;
    const int c = static_cast<int>(i - pix * C);
    const float v = x[pix * Ct + c0 + c];
    const float h = round_tf32(v);
    hi[i] = h;
    lo[i] = round_tf32(v - h);
  }
}

void launch_split_tf32(const float* x, int Ct, int c0, int C, float* hi, float* lo, size_t npix, cudaStream_t st) {
  const size_t n = npix * static_cast<size_t>(C);
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 32));
  split_tf32_kernel<<<blocks, 256, 0, st>>>(x, Ct, c0, C, hi, lo, npix);
}

// ------------------------------------------------------------------------------------------------
// upfirdn2d, the reference's own native op (op/upfirdn2d.cpp:12-23, semantics of upfirdn2d_native,
// op/upfirdn2d.py:173-208): zero-insert upsample, pad (negative pads crop), FIR with the flipped kernel,
// decimate.  Generic gather formulation over [major][H][W][minor]; the network itself uses the fused
// normalise+FIR kernel above, this entry point exists for callers of the reference's FFI seam.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upfirdn2d_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         const float* __restrict__ kernel, int in_h, int in_w, int minor,
                                                         int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                                                         int pad_x0, int pad_y0, int out_h, int out_w, long long total) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int mi = static_cast<int>(i % minor);
    const int ox = static_cast<int>((i / minor) % out_w);
    const int oy = static_cast<int>((i / minor / out_w) % out_h);
    const long long mj = i / minor / out_w / out_h;
    float acc = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int uy = oy * down_y + ky - pad_y0;
      if (uy < 0 || uy >= in_h * up_y || uy % up_y) continue;
      const int iy = uy / up_y;
      for (int kx = 0; kx < kw; ++kx) {
        const int ux = ox * down_x + kx - pad_x0;
        if (ux < 0 || ux >= in_w * up_x || ux % up_x) continue;
        const int ix = ux / up_x;
        acc += in[((mj * in_h + iy) * in_w + ix) * minor + mi] * kernel[(kh - 1 - ky) * kw + (kw - 1 - kx)];
      }
    }
    out[i] = acc;
  }
}

void launch_upfirdn2d(const float* in, float* out, int major, int in_h, int in_w, int minor, const float* kernel, int kh,
                      int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                      cudaStream_t st) {
  const int out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) / down_y + 1;
  const int out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) / down_x + 1;
  const long long total = static_cast<long long>(major) * out_h * out_w * minor;
  if (total <= 0) return;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 32));
  upfirdn2d_kernel<<<blocks, 256, 0, st>>>(in, out, kernel, in_h, in_w, minor, kh, kw, up_x, up_y, down_x, down_y, pad_x0,
                                           pad_y0, out_h, out_w, total);
}

}  // namespace use
