// tcgen05 implicit-GEMM convolution (3x3 pad 1 / 1x1, stride 1) over NHWC activations.
//
// Replaces, for the hot path, every nn.Conv2d with C_in, C_out >= 64 that the reference runs through
// cuDNN/oneDNN (layers.py:113-162 called from layerspp.py:282-314; 94.7 % + 5 % of the FLOPs,
// SURVEY.md section 3.3).  One launch can sum several "segments" into the same accumulator:
//   Conv_1(3x3 over a1) + Conv_2(1x1 over raw x [+ 1x1 over the skip tensor])   (layerspp.py:306-309)
// so the skip projection and the residual add cost no extra pass over HBM.
//
// GEMM view: D[M = 128 pixels][N = C_out] += A[pixels][K = 128 B of channels] * W[C_out][K]^T per filter tap.
// Tile = 8 (w) x 16*NSUB (h) output pixels.  Per channel chunk and horizontal tap s the TMA producer
// loads ONE (16*NSUB+2) x 8 pixel box (zero-filled outside the image = the conv padding); each image row
// of the box is exactly one 1024-byte SWIZZLE_128B atom, so the three vertical taps r are the same smem
// tile at +r*1024 B: 3 activation loads per chunk instead of 9.  NSUB sub-tiles share every weight tile.
//
// Warp roles: warp 0 = TMA producer (1 thread), warp 1 = TMEM owner + MMA issuer (1 thread),
// warps 2.. = epilogue (TMEM -> registers -> bias / residual / scale -> global).  Accumulators are
// double-buffered in TMEM (2 x NSUB x N columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Persistent CTAs, static round-robin over tiles (w fastest: neighbours share halos and weights in L2).
#pragma once
#include "common.cuh"

namespace use {

struct alignas(64) ConvSeg {
  CUtensorMap tmA;  // rank 4 {C, W, H, B}, box {CK, 8, rows, 1}; rows = 16*NSUB+2 (3x3) or 16*NSUB (1x1)
  CUtensorMap tmW;  // rank 3 {C_total, N, taps}, box {CK, N, 1}
  int nchunks;      // channels of this segment / CK
  int taps;         // 9 or 1
  int wc0;          // first weight channel of this segment inside tmW (concatenated inputs)
  int ac0;          // first channel inside tmA
};

struct alignas(64) ConvParams {
  ConvSeg seg[3];
  int nseg;
  int B, H, W;
  int tiles_w, tiles_h, ntiles;
  void* out;          // T [B][H][W][N]
  const float* bias;  // [B or 1][N]
  int bias_bstride;   // N (per-sample bias incl. the time-embedding term) or 0
  const void* res;    // optional residual, T [B][H][W][N]
  float scale;        // out = (acc + bias [+ res]) * scale
};

template <typename T, int N, int NSUB>
struct ConvCfg {
  static constexpr int CK = 128 / sizeof(T);
  static constexpr int TILE_W = 8;
  static constexpr int TILE_H = 16 * NSUB;
  static constexpr int A_ROWS = TILE_H + 2;
  static constexpr int A_SLOT = A_ROWS * 1024;
  static constexpr int B_TILE = N * 128;
  static constexpr int A_SLOTS = 3;
  static constexpr int B_SLOTS = (N == 256) ? 5 : (N == 128 ? 7 : 8);
  static constexpr int ACC_COLS = NSUB * N;
  static constexpr int TMEM_COLS = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;
  static constexpr int EPI_WARPS = 4 * NSUB;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int NBARS = 2 * A_SLOTS + 2 * B_SLOTS + 4;
  static constexpr int SMEM_BYTES = 1024 + A_SLOTS * A_SLOT + B_SLOTS * B_TILE + NBARS * 8 + 16;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of two <= 512");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(N % 32 == 0 && N <= 256, "N");
};

template <typename T, int N, int NSUB>
__global__ void __launch_bounds__(ConvCfg<T, N, NSUB>::THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  using C = ConvCfg<T, N, NSUB>;
  constexpr bool kBf16 = DT<T>::kIsBf16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + C::A_SLOTS * C::A_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C::B_SLOTS * C::B_TILE);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + C::A_SLOTS;
  uint64_t* b_full = a_empty + C::A_SLOTS;
  uint64_t* b_empty = b_full + C::B_SLOTS;
  uint64_t* t_full = b_empty + C::B_SLOTS;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nseg; ++i) {
      prefetch_tmap(&p.seg[i].tmA);
      prefetch_tmap(&p.seg[i].tmW);
    }
    for (int i = 0; i < C::A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < C::B_SLOTS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], C::EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      uint32_t ai = 0, bi = 0;  // running slot counters
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img;
        const int rem = tile - b * tiles_per_img;
        const int th = rem / p.tiles_w;
        const int w0 = (rem - th * p.tiles_w) * C::TILE_W;
        const int h0 = th * C::TILE_H;
        for (int sg = 0; sg < p.nseg; ++sg) {
          const ConvSeg& S = p.seg[sg];
          const bool k3 = S.taps == 9;
          const int ns = k3 ? 3 : 1;
          const uint32_t a_bytes = (k3 ? C::A_ROWS : C::TILE_H) * 1024;
          for (int kc = 0; kc < S.nchunks; ++kc) {
            for (int s = 0; s < ns; ++s) {
              const uint32_t as = ai % C::A_SLOTS, aph = (ai / C::A_SLOTS) & 1;
              mbar_wait(&a_empty[as], aph ^ 1);
              mbar_arrive_expect_tx(&a_full[as], a_bytes);
              tma_load_4d(sA + as * C::A_SLOT, &S.tmA, &a_full[as], S.ac0 + kc * C::CK, k3 ? (w0 + s - 1) : w0,
                          k3 ? (h0 - 1) : h0, b);
              ++ai;
              for (int r = 0; r < ns; ++r) {
                const uint32_t bs = bi % C::B_SLOTS, bph = (bi / C::B_SLOTS) & 1;
                mbar_wait(&b_empty[bs], bph ^ 1);
                mbar_arrive_expect_tx(&b_full[bs], C::B_TILE);
                tma_load_3d(sB + bs * C::B_TILE, &S.tmW, &b_full[bs], S.wc0 + kc * C::CK, 0, k3 ? (r * 3 + s) : 0);
                ++bi;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kBf16 ? 1 : 2, 128, N);
      const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
      uint32_t ai = 0, bi = 0, ti = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
        const uint32_t acs = ti & 1, acph = (ti >> 1) & 1;
        mbar_wait(&t_empty[acs], acph ^ 1);
        tc_fence_after();
        bool first = true;
        for (int sg = 0; sg < p.nseg; ++sg) {
          const ConvSeg& S = p.seg[sg];
          const int ns = S.taps == 9 ? 3 : 1;
          for (int kc = 0; kc < S.nchunks; ++kc) {
            for (int s = 0; s < ns; ++s) {
              const uint32_t as = ai % C::A_SLOTS, aph = (ai / C::A_SLOTS) & 1;
              mbar_wait(&a_full[as], aph);
              for (int r = 0; r < ns; ++r) {
                const uint32_t bs = bi % C::B_SLOTS, bph = (bi / C::B_SLOTS) & 1;
                mbar_wait(&b_full[bs], bph);
                tc_fence_after();
#pragma unroll
                for (int sub = 0; sub < NSUB; ++sub) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = umma_desc_sw128(sA_addr + as * C::A_SLOT + (sub * 16 + r) * 1024 + k * 32);
                    const uint64_t bd = umma_desc_sw128(sB_addr + bs * C::B_TILE + k * 32);
                    umma_ss<kBf16>(tmem_base + acs * C::ACC_COLS + sub * N, ad, bd, idesc, (first && k == 0) ? 0u : 1u);
                  }
                }
                first = false;
                umma_commit(&b_empty[bs]);
                ++bi;
              }
              umma_commit(&a_empty[as]);
              ++ai;
            }
          }
        }
        umma_commit(&t_full[acs]);
      }
    }
  } else {
    // ================================ epilogue ================================
    const int ew = warp - 2;
    const int sub = ew >> 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int m = quad * 32 + lane;
    const int hl = sub * 16 + (m >> 3), wl = m & 7;
    T* out = reinterpret_cast<T*>(p.out);
    const T* res = reinterpret_cast<const T*>(p.res);
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
      const int b = tile / tiles_per_img;
      const int rem = tile - b * tiles_per_img;
      const int th = rem / p.tiles_w;
      const int w = (rem - th * p.tiles_w) * C::TILE_W + wl;
      const int h = th * C::TILE_H + hl;
      const bool valid = (h < p.H) && (w < p.W);
      const size_t pix = (static_cast<size_t>(b) * p.H + h) * p.W + w;
      const float* bias = p.bias + static_cast<size_t>(b) * p.bias_bstride;
      const uint32_t acs = ti & 1, acph = (ti >> 1) & 1;
      mbar_wait(&t_full[acs], acph);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acs * C::ACC_COLS + sub * N;
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(trow + c0, r);
        tmem_ld_wait();
        if (valid) {
          constexpr int V = DT<T>::kVec;
#pragma unroll
          for (int j = 0; j < 32; j += V) {
            float v[V];
#pragma unroll
            for (int q = 0; q < V; q += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + j + q));
              v[q] = __uint_as_float(r[j + q]) + bb.x;
              v[q + 1] = __uint_as_float(r[j + q + 1]) + bb.y;
              v[q + 2] = __uint_as_float(r[j + q + 2]) + bb.z;
              v[q + 3] = __uint_as_float(r[j + q + 3]) + bb.w;
            }
            if (res != nullptr) {
              float rr[V];
              Vec<T>::load(res + pix * N + c0 + j, rr);
#pragma unroll
              for (int q = 0; q < V; ++q) v[q] += rr[q];
            }
#pragma unroll
            for (int q = 0; q < V; ++q) v[q] *= p.scale;
            Vec<T>::store(out + pix * N + c0 + j, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acs]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace use
