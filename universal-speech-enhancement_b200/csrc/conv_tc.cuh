// tcgen05 implicit-GEMM convolution (3x3 pad 1 / 1x1, stride 1) over NHWC activations.
//
// Replaces, for the hot path, every nn.Conv2d with C_in, C_out >= 64 that the reference runs through
// cuDNN/oneDNN (layers.py:113-162 called from layerspp.py:282-314; 94.7 % + 5 % of the FLOPs,
// SURVEY.md section 3.3).  One launch can sum several "segments" into the same accumulator:
//   Conv_1(3x3 over a1) + Conv_2(1x1 over raw x [+ 1x1 over the skip tensor])   (layerspp.py:306-309)
// so the skip projection and the residual add cost no extra pass over HBM.
//
// GEMM view: D[M = 128 pixels][N = C_out] += A[pixels][K = 128 B of channels] * W[C_out][K]^T per filter tap.
// Tile = 8 (w) x 16*NSUB (h) output pixels.  Per channel chunk ONE (16*NSUB+2) x 10 pixel window is staged in shared
// memory, 128 B per pixel, pixels 128 B apart, every 128-byte row swizzled by its absolute address (SWIZZLE_128B).
// All nine filter taps read that one window: the operand descriptor of tap (r, s) simply starts at pixel
// (r, s) of the window and steps SBO = 1280 B (one window row) between 8-pixel groups -- the hardware swizzle is a
// function of the absolute shared-memory address (measured: tools/swz_probe.cu), so a descriptor may start at any
// 128-byte row.  1 activation stage per chunk instead of 9 (or 3 horizontally shifted copies).
//
// The window is always landed by TMA (zero fill outside the image = the conv padding).  For a FUSED segment the TMA
// brings the RAW producer output and six "transform" warps apply GroupNorm (per-sample, per-channel scale / shift
// table) + SiLU + operand rounding IN PLACE in shared memory before the MMAs read it: the normalised activation tensor
// of `GroupNorm -> SiLU -> Conv3x3` (layerspp.py:283-285,304-306) is never materialised in HBM.  (Measured
// alternatives: global -> registers -> smem in the transform warps exposes the load latency, 70 % of the plain conv
// speed; in place behind a 3-deep TMA ring: 92-94 %.  L2 prefetch hints made both variants slower.)
//
// Warp roles: warp 0 = TMA producer (1 thread), warp 1 = TMEM owner + MMA issuer (the warp in lockstep, one elected lane),
// warps 2..2+4*NSUB = epilogue (TMEM -> registers -> bias / residual / scale -> global), then 6 transform warps.
// Accumulators are double-buffered in TMEM (2 x NSUB x N columns) so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Persistent CTAs, static round-robin over tiles (w fastest: neighbours share halos and weights in L2).
#pragma once
#include "common.cuh"

namespace use {

struct alignas(64) ConvSeg {
  CUtensorMap tmA;  // rank 4 {C, W, H, B}; 3x3: box {CK, 10, 16*NSUB+2, 1}; 1x1: box {CK, 8, 16*NSUB, 1}
  CUtensorMap tmW;  // rank 3 {C_total, N, taps}, box {CK, N, 1}
  CUtensorMap tmWh; // same tensor, box {CK, N/2, 1}: the half each CTA of a pair loads and multicasts
  int nchunks;      // channels of this segment / CK
  int taps;         // 9 or 1
  int wc0;          // first weight channel of this segment inside tmW (concatenated inputs)
  int ac0;          // first channel inside the activation tensor
  // fused GroupNorm + SiLU operand (3x3 only): raw != nullptr -> the transform warps fill the window
  const void* raw;   // T [B][H][W][Ct]
  const float* aff;  // fp32 [B][2][aff_C]: per-sample scale row, then shift row (launch_gn_affine)
  int Ct;            // channel pitch of raw
  int aff_C;         // row length of aff
  int aff_c0;        // channel of aff that corresponds to channel ac0 of raw
};

struct alignas(64) ConvParams {
  ConvSeg seg[3];
  int nseg;
  int B, H, W;
  int tiles_w, tiles_h, ntiles;
  // C_out split ("N-split") for launches with fewer tiles than SMs (small batches at the deep levels, where ONE CTA
  // would stream a whole layer's weights through one SM: 29 us per 8 x 10 layer): the kernel's N is a 64-channel SLICE
  // of the ldn = C_out channels, a work unit is (tile, slice) and nunits = ntiles * nsplit of them are spread over the
  // CTAs.  Every output element sees the same K sequence as in the unsplit kernel: results are bit-identical.
  int nsplit, ldn, nunits;
  int tile_base;  // first tile of this launch (a layer can be run as a balanced main launch + a sliced tail launch);
                  // ntiles is the END of this launch's tile range, nunits = (ntiles - tile_base) * nsplit
  // Inline GroupNorm (fused segments only; gn_st0 != nullptr): the scale / shift table of GroupNorm(cat[s0, s1]) is NOT read
  // from global memory (seg[].aff, one gn_affine_kernel launch per convolution: 98 launches of ~6 us per evaluation) but
  // computed by the transform warps into shared memory whenever the CTA moves on to another sample -- the arithmetic of
  // gn_affine_kernel, bit for bit.  Channel c of the concatenation = channel c of s0 (c < gn_C0) or c - gn_C0 of s1.
  const long long* gn_st0;
  const long long* gn_st1;
  const float* gn_gamma;
  const float* gn_beta;
  int gn_C0, gn_C1, gn_HW;
  float gn_eps;
  void* out;          // T [B][H][W][ldn]
  const float* bias;  // [B or 1][N]
  int bias_bstride;   // N (per-sample bias incl. the time-embedding term) or 0
  const void* res;    // optional residual, T [B][H][W][N]
  float scale;        // out = (acc + bias [+ res]) * scale
  long long* stats_acc;  // optional [B][N][2] fixed-point accumulators (zero on entry): sum / sum of squares of `out`
  unsigned long long* prof;  // PROF kernels only: per-role stall counters (tools/conv_bench.py --prof)
  int dbg;                   // PROF kernels only (USE_B200_CONV_DBG): 1 = epilogue skips TMEM loads and stores, 2 = epilogue
                             // loads TMEM but stores nothing, 4 = transform skips its in-place pass
};

// PROF instrumentation: cycles a role's elected thread spends inside an mbarrier wait
template <bool PROF>
__device__ __forceinline__ void mbar_wait_p(uint64_t* bar, uint32_t parity, long long& acc) {
  if constexpr (PROF) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}

template <typename T, int N, int NSUB, bool FUSE, bool CG2 = false>
struct ConvCfg {
  static constexpr int CK = 128 / sizeof(T);
  static constexpr int TILE_W = 8;
  static constexpr int TILE_H = 16 * NSUB;
  static constexpr int A_ROWS = TILE_H + 2;
  static constexpr int WIN_W = 10;                       // window width in pixels (8 + halo)
  static constexpr int WIN_PITCH = WIN_W * 128;          // bytes between window rows = SBO of a 3x3 operand
  static constexpr int NPIX = A_ROWS * WIN_W;            // pixels of one window
  static constexpr int A_SLOT = (NPIX * 128 + 1023) & ~1023;
  static constexpr int B_TILE = CG2 ? N * 64 : N * 128;  // CTA pair: each CTA stages half of the C_out rows
  // plain: 2 windows (loading / consumed).  fused: 3 (TMA loading the raw window / being normalised in place / consumed)
  static constexpr int A_SLOTS = FUSE ? 3 : 2;
  // (N <= 64: 8 slots of 8 KB.  16 slots for the N-split slice kernel were measured: batch 1 158.6 vs 158.3 ms -- no gain)
  static constexpr int B_SLOTS = CG2 ? (FUSE ? 10 : 16)
                                     : (N == 256) ? (FUSE ? 4 : 5) : (N == 128 ? (FUSE ? 5 : 8) : 8);
  static constexpr int ACC_COLS = NSUB * N;
  static constexpr int TMEM_COLS = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;
  static constexpr int EPI_WARPS = 4 * NSUB;
  static constexpr int XF_THREADS = FUSE ? 192 : 0;      // transform warps (6)
  static constexpr int XF_T0 = 64 + 32 * EPI_WARPS;      // first transform thread
  static constexpr int THREADS = XF_T0 + XF_THREADS;
  static constexpr int NBARS = 3 * A_SLOTS + 2 * B_SLOTS + 4;
  static constexpr int STAT_BYTES = EPI_WARPS * N * 2 * 4;  // per-warp column statistics of the current tile
  static constexpr int GN_MAXC = 512;                       // channels of an inline GroupNorm (cat of two 256-channel tensors)
  static constexpr int GN_BYTES = FUSE ? 2 * GN_MAXC * 4 : 0;  // scale row, shift row
  static constexpr int SMEM_BYTES = 1024 + A_SLOTS * A_SLOT + B_SLOTS * B_TILE + STAT_BYTES + GN_BYTES + NBARS * 8 + 16;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of two <= 512");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(N % 32 == 0 && N <= 256, "N");
};

// SWAP (C_out = 128 only): "swap-AB".  A single-CTA M128 x N128 MMA reads 4 KB (A) + 4 KB (B) of shared memory per 64
// cycles = the 128 B/clk shared-memory limit, which capped these layers (80 % of the FLOPs) near 70 % of the tensor
// peak.  With the roles swapped -- weights [128 c_out x K] as the M operand, the 256 pixels of the tile as the N
// operand -- one MMA reads 4 + 8 KB per 128 cycles (96 B/clk, the profile of the C_out = 256 layers).  The accumulator
// is then channel-major (TMEM lane = output channel, column = pixel); the epilogue stores 32 consecutive channels per
// pixel per warp, and per-channel GroupNorm statistics become in-register sums.
//
// MC = 2: CTA pairs (cluster of 2).  Every tile re-reads the layer's whole weight tensor (295 KB for 128 -> 128 in bf16)
// from L2, and with 148 CTAs doing so the L2 -> SM weight stream is what the MMA issuer waits for (measured with the
// PROF build: 25 % of its time in b_full waits, 1 % in a_full).  Both CTAs of a pair walk their own tiles through the
// same weight-tile sequence; each loads HALF of every weight tile and multicasts it into both CTAs' shared memory,
// which halves that stream.  A weight slot is recycled only when the MMAs of BOTH CTAs have released it (b_empty
// counts 2 arrivals, one of them a multicast tcgen05.commit from the peer).  A CTA whose tile index falls past the end
// processes a "ghost" tile (loads are zero-filled out of bounds, nothing is stored) to keep the pair in lockstep.
// Measured: no gain (a multicast to fewer than ~8 CTAs is not deduplicated in L2 on this part); off by default.
//
// CG2 (C_out = 128, pixel-major, MC = 2): the CTA pair issues ONE tcgen05.mma.cta_group::2 per step (M = 256 = the 128
// pixels of a sub-tile of each CTA, N = 128): each CTA stages only HALF of every weight tile (its 64 C_out rows; the
// tensor cores exchange the halves), which halves the L2 -> SM weight stream per SM -- the measured limiter of these
// layers -- and doubles the time the weight ring covers.  Only the leader (cluster rank 0) issues MMAs; every barrier
// that collects both CTAs' arrivals (a_full, b_full, t_empty) lives in the leader and is signalled remotely by the
// peer's TMA loads (cp.async.bulk.tensor.cta_group::2), transform warps and epilogue warps; the leader's
// tcgen05.commit releases slots / publishes accumulators in both CTAs by multicast.  Measured: bit-correct, but slower
// than swap-AB (1332 vs 1713 TFLOP/s, bf16 128 -> 128): the leader issues eight N = 128 MMAs per tap for two SMs and the
// pixel-major epilogue is heavier; USE_B200_CONV_CG2=1 selects it for experiments.
template <typename T, int N, int NSUB, bool SWAP, bool FUSE, int MC = 1, bool PROF = false, bool CG2 = false>
__global__ void __launch_bounds__(ConvCfg<T, N, NSUB, FUSE, CG2>::THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  using C = ConvCfg<T, N, NSUB, FUSE, CG2>;
  static_assert(!SWAP || (N == 128 && NSUB == 2), "swap-AB is built for C_out = 128, 256-pixel tiles");
  static_assert(!CG2 || (!SWAP && MC == 2 && N == 128 && NSUB == 2), "the CTA-pair MMA form is built for C_out = 128");
  constexpr bool kBf16 = DT<T>::kIsBf16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + C::A_SLOTS * C::A_SLOT;
  float* stat_s = reinterpret_cast<float*>(sB + C::B_SLOTS * C::B_TILE);
  float* gn_s = reinterpret_cast<float*>(sB + C::B_SLOTS * C::B_TILE + C::STAT_BYTES);  // [2][GN_MAXC] (FUSE)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C::B_SLOTS * C::B_TILE + C::STAT_BYTES + C::GN_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + C::A_SLOTS;
  uint64_t* a_raw = a_empty + C::A_SLOTS;  // fused segments: raw window landed (TMA -> transform warps)
  uint64_t* b_full = a_raw + C::A_SLOTS;
  uint64_t* b_empty = b_full + C::B_SLOTS;
  uint64_t* t_full = b_empty + C::B_SLOTS;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // PDL: let the next kernel of the stream be scheduled now (its CTAs only do local set-up before their own pdl_wait);
  // everything up to pdl_wait() below touches nothing but kernel parameters, shared memory and TMEM
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nseg; ++i) {
      prefetch_tmap(&p.seg[i].tmA);
      prefetch_tmap(&p.seg[i].tmW);
      prefetch_tmap(&p.seg[i].tmWh);
    }
    // CG2: the "full" barriers and t_empty of the LEADER collect one arrival (or one group of arrivals) from each CTA
    constexpr int kPair = CG2 ? 2 : 1;
    for (int i = 0; i < C::A_SLOTS; ++i) { mbar_init(&a_full[i], kPair); mbar_init(&a_empty[i], 1); mbar_init(&a_raw[i], 1); }
    for (int i = 0; i < C::B_SLOTS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], CG2 ? 1 : MC); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], kPair * C::EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG2) { tmem_alloc_cg2(tmem_slot, C::TMEM_COLS); tmem_relinquish_cg2(); }
    else { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC > 1) cluster_sync_all();  // the peer's barriers are initialised before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the predecessor's outputs (activations, statistics, scale / shift tables) are complete from here on

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  // tile groups of MC consecutive tiles: CTA `crank` of a pair takes tile group * MC + crank.  T0 / TSTEP / TEND walk the
  // tile index of THIS CTA; TEND is rounded up so both CTAs of a pair run the same number of iterations.
  const int crank = MC > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  // (with an N-split the walk is over work units u = tile * nsplit + slice; nsplit = 1: u = tile)
  const int T0 = (blockIdx.x / MC) * MC + crank, TSTEP = (gridDim.x / MC) * MC;
  const int TEND = (p.nunits + MC - 1) / MC * MC;
  const int nsp = p.nsplit;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // Two streams issued by one thread: weight tiles (B ring) and activation windows (A ring).  The window stream runs
    // ahead of the weight stream by up to A_SLOTS - 1 chunks (it is polled, non-blocking, before every weight tile), so a
    // raw window of a fused segment lands -- and is normalised in place by the transform warps -- well before its MMAs.
    if (lane == 0) {
      // window-stream cursor
      int a_tile = T0, a_sg = 0, a_kc = 0;
      uint32_t ai = 0;
      long long w_a = 0, w_b = 0;
      const long long t_begin = PROF ? clock64() : 0;
      auto a_pending = [&]() { return a_tile < TEND; };
      auto a_issue = [&](bool blocking) -> bool {
        const uint32_t as = ai % C::A_SLOTS, aph = (ai / C::A_SLOTS) & 1;
        if (blocking) mbar_wait_p<PROF>(&a_empty[as], aph ^ 1, w_a);
        else if (!mbar_test_wait(&a_empty[as], aph ^ 1)) return false;
        const ConvSeg& S = p.seg[a_sg];
        const int a_t = p.tile_base + a_tile / nsp;  // a_tile walks work units
        const int b = a_t / tiles_per_img;
        const int rem = a_t - b * tiles_per_img;
        const int th = rem / p.tiles_w;
        const int w0 = (rem - th * p.tiles_w) * C::TILE_W, h0 = th * C::TILE_H;
        const bool k3 = S.taps == 9;
        const uint32_t a_bytes = (k3 ? C::NPIX : C::TILE_H * 8) * 128;
        if (FUSE && S.raw != nullptr) {
          mbar_arrive_expect_tx(&a_raw[as], a_bytes);
          tma_load_4d(sA + as * C::A_SLOT, &S.tmA, &a_raw[as], S.ac0 + a_kc * C::CK, k3 ? (w0 - 1) : w0, k3 ? (h0 - 1) : h0, b);
        } else if constexpr (CG2) {
          // the leader's barrier takes the bytes of BOTH CTAs' windows.  No remote arrive here: a cluster-scope arrive
          // costs this thread ~1 us, and it also has to keep the weight stream going (measured: 2.5x slower with it).
          // a_full counts 2 (a fused window is published by one transform thread per CTA): the leader arrives twice.
          const uint32_t lbar = mapa_u32(smem_u32(&a_full[as]), 0);
          if (crank == 0) {
            mbar_arrive_expect_tx(&a_full[as], 2 * a_bytes);
            mbar_arrive(&a_full[as]);
          }
          tma_load_4d_cg2(sA + as * C::A_SLOT, &S.tmA, lbar, S.ac0 + a_kc * C::CK, k3 ? (w0 - 1) : w0, k3 ? (h0 - 1) : h0, b);
        } else {
          mbar_arrive_expect_tx(&a_full[as], a_bytes);
          tma_load_4d(sA + as * C::A_SLOT, &S.tmA, &a_full[as], S.ac0 + a_kc * C::CK, k3 ? (w0 - 1) : w0, k3 ? (h0 - 1) : h0, b);
        }
        ++ai;
        if (++a_kc == S.nchunks) {
          a_kc = 0;
          if (++a_sg == p.nseg) { a_sg = 0; a_tile += TSTEP; }
        }
        return true;
      };
      uint32_t bi = 0, bj = 0;  // weight tiles issued, chunks whose weights have been issued
      for (int tile = T0; tile < TEND; tile += TSTEP) {
        const int wrow0 = (tile % nsp) * N;  // first C_out row of this unit's slice
        for (int sg = 0; sg < p.nseg; ++sg) {
          const ConvSeg& S = p.seg[sg];
          const bool k3 = S.taps == 9;
          for (int kc = 0; kc < S.nchunks; ++kc, ++bj) {
            while (ai <= bj) a_issue(true);  // the window of this chunk is always issued before its weights
            for (int tap = 0; tap < S.taps; ++tap) {
              if (a_pending() && ai < bj + C::A_SLOTS) a_issue(false);
              if constexpr (PROF) { if (p.dbg & 8) continue; }
              const uint32_t bs = bi % C::B_SLOTS, bph = (bi / C::B_SLOTS) & 1;
              mbar_wait_p<PROF>(&b_empty[bs], bph ^ 1, w_b);
              // tap order s-major (s = tap / 3, r = tap % 3): the accumulation order of the previous 3-copy kernel
              const int wtap = k3 ? ((tap % 3) * 3 + tap / 3) : 0;
              if constexpr (CG2) {
                const uint32_t lbar = mapa_u32(smem_u32(&b_full[bs]), 0);
                if (crank == 0) mbar_arrive_expect_tx(&b_full[bs], 2 * C::B_TILE);  // both halves land on the leader's barrier
                tma_load_3d_cg2(sB + bs * C::B_TILE, &S.tmWh, lbar, S.wc0 + kc * C::CK, crank * (N / 2), wtap);
                ++bi;
                continue;
              }
              mbar_arrive_expect_tx(&b_full[bs], C::B_TILE);
              if constexpr (MC > 1)
                tma_load_3d_mc(sB + bs * C::B_TILE + crank * (C::B_TILE / 2), &S.tmWh, &b_full[bs], S.wc0 + kc * C::CK,
                               crank * (N / 2), wtap, uint16_t(3));
              else
                tma_load_3d(sB + bs * C::B_TILE, &S.tmW, &b_full[bs], S.wc0 + kc * C::CK, wrow0, wtap);
              ++bi;
            }
          }
        }
      }
      if constexpr (PROF) {
        atomicAdd(p.prof + 4, static_cast<unsigned long long>(w_a));
        atomicAdd(p.prof + 5, static_cast<unsigned long long>(w_b));
        atomicAdd(p.prof + 6, static_cast<unsigned long long>(clock64() - t_begin));
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The whole warp runs this loop in lockstep and one ELECTED lane issues the tcgen05 instructions: with uniform control
    // flow the descriptors stay in uniform registers.  (Issued from inside an `if (lane == 0)` the compiler wraps every
    // tcgen05.mma in a register-broadcast loop, ~100 cycles per instruction -- enough to starve the CTA-pair form, whose
    // leader issues for two SMs.)
    if (!CG2 || crank == 0) {  // CTA pair: only the leader issues (for both CTAs)
      constexpr uint32_t idesc = SWAP ? umma_idesc(kBf16 ? 1 : 2, 128, 128 * NSUB)
                                      : umma_idesc(kBf16 ? 1 : 2, CG2 ? 256 : 128, N);
      const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
      uint32_t ai = 0, bi = 0, ti = 0;
      long long w_t = 0, w_a = 0, w_b = 0;
      const long long t_begin = PROF ? clock64() : 0;
      for (int tile = T0; tile < TEND; tile += TSTEP, ++ti) {
        const uint32_t acs = ti & 1, acph = (ti >> 1) & 1;
        mbar_wait_p<PROF>(&t_empty[acs], acph ^ 1, w_t);
        tc_fence_after();
        bool first = true;
        for (int sg = 0; sg < p.nseg; ++sg) {
          const ConvSeg& S = p.seg[sg];
          const bool k3 = S.taps == 9;
          const uint32_t sbo = k3 ? C::WIN_PITCH : 1024;
          for (int kc = 0; kc < S.nchunks; ++kc) {
            const uint32_t as = ai % C::A_SLOTS, aph = (ai / C::A_SLOTS) & 1;
            mbar_wait_p<PROF>(&a_full[as], aph, w_a);
            for (int tap = 0; tap < S.taps; ++tap) {
              const int s = tap / 3, r = tap - s * 3;  // (0, 0) for a 1x1 segment
              const uint32_t bs = bi % C::B_SLOTS, bph = (bi / C::B_SLOTS) & 1;
              bool dbg_nob = false;
              if constexpr (PROF) dbg_nob = (p.dbg & 8) != 0;
              if (!dbg_nob) mbar_wait_p<PROF>(&b_full[bs], bph, w_b);
              tc_fence_after();
              const uint32_t win = sA_addr + as * C::A_SLOT + (k3 ? (r * C::WIN_W + s) * 128 : 0);
              if (elect_one()) {
              if constexpr (SWAP) {
                // D[c_out][pixel] += W[c_out][K] * X[pixel][K]^T : M operand = weight tile, N operand = 256 pixel rows
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t wd = umma_desc_sw128(sB_addr + bs * C::B_TILE + k * 32);
                  const uint64_t xd = umma_desc_sw128_sbo(win + k * 32, sbo);
                  umma_ss<kBf16>(tmem_base + acs * C::ACC_COLS, wd, xd, idesc, (first && k == 0) ? 0u : 1u);
                }
              } else {
#pragma unroll
                for (int sub = 0; sub < NSUB; ++sub) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = umma_desc_sw128_sbo(win + sub * 16 * sbo + k * 32, sbo);
                    const uint64_t bd = umma_desc_sw128(sB_addr + bs * C::B_TILE + k * 32);
                    if constexpr (CG2)
                      umma_ss_cg2<kBf16>(tmem_base + acs * C::ACC_COLS + sub * N, ad, bd, idesc, (first && k == 0) ? 0u : 1u);
                    else
                      umma_ss<kBf16>(tmem_base + acs * C::ACC_COLS + sub * N, ad, bd, idesc, (first && k == 0) ? 0u : 1u);
                  }
                }
              }
              if constexpr (CG2) umma_commit_cg2(&b_empty[bs], uint16_t(3));
              else if (dbg_nob) {
              } else if constexpr (MC > 1) umma_commit_mc(&b_empty[bs], uint16_t(3));
              else umma_commit(&b_empty[bs]);
              }  // elected lane
              __syncwarp();
              first = false;
              ++bi;
            }
            if (elect_one()) {
              if constexpr (CG2) umma_commit_cg2(&a_empty[as], uint16_t(3));
              else umma_commit(&a_empty[as]);
            }
            __syncwarp();
            ++ai;
          }
        }
        if (elect_one()) {
          if constexpr (CG2) umma_commit_cg2(&t_full[acs], uint16_t(3));
          else umma_commit(&t_full[acs]);
        }
        __syncwarp();
      }
      if constexpr (PROF) if (lane == 0) {
        atomicAdd(p.prof + 0, static_cast<unsigned long long>(w_t));
        atomicAdd(p.prof + 1, static_cast<unsigned long long>(w_a));
        atomicAdd(p.prof + 2, static_cast<unsigned long long>(w_b));
        atomicAdd(p.prof + 3, static_cast<unsigned long long>(clock64() - t_begin));
      }
    }
  } else if (threadIdx.x >= C::XF_T0) {
    if constexpr (FUSE) {
    // ================================ transform warps ================================
    // GroupNorm (scale / shift) + SiLU + operand rounding of the raw window the TMA just landed, IN PLACE in shared
    // memory (the transform never waits on global memory).  Thread = one 16-byte channel vector (v) of the window
    // pixels pb, pb + PSTEP, ...; pixels outside the image keep the TMA's zero fill: the conv pads the ACTIVATED tensor.
    constexpr int V = DT<T>::kVec;
    constexpr int XT = C::XF_THREADS;
    constexpr int PSTEP = XT / 8;
    constexpr int NIT = (C::NPIX + PSTEP - 1) / PSTEP;
    const int tt = threadIdx.x - C::XF_T0;
    const int v = tt & 7, pb = tt >> 3;
    uint32_t ai = 0, rawph = 0;  // rawph: phase parity of a_raw per slot (it only advances on fused fills)
    int gn_b = -1;               // sample whose inline GroupNorm table is in gn_s
    long long w_r = 0;
    const long long t_begin = PROF ? clock64() : 0;
    for (int unit = T0; unit < TEND; unit += TSTEP) {
      const int tile = p.tile_base + unit / nsp;
      const int b = tile / tiles_per_img;
      const int rem = tile - b * tiles_per_img;
      const int th = rem / p.tiles_w;
      const int w0 = (rem - th * p.tiles_w) * C::TILE_W;
      const int h0 = th * C::TILE_H;
      uint32_t inside = 0;  // bit i: window pixel pb + PSTEP * i lies inside the image
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const int q = pb + PSTEP * i;
        const int row = q / C::WIN_W, col = q - row * C::WIN_W;
        const int hh = h0 - 1 + row, ww = w0 - 1 + col;
        if (tile < p.ntiles && q < C::NPIX && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) inside |= 1u << i;
      }
      if (p.gn_st0 != nullptr && tile < p.ntiles && b != gn_b) {
        // a new sample: rebuild the scale / shift table (every transform thread is past the previous tile's last chunk
        // barrier, so nobody still reads the old table).  Same arithmetic as gn_affine_kernel.
        const int Ct = p.gn_C0 + p.gn_C1;
        const int G = min(Ct / 4, 32), cpg = Ct / G;
        const double inv_cnt = 1.0 / (static_cast<double>(p.gn_HW) * cpg);
        for (int c = tt; c < Ct; c += XT) {
          const int g = c / cpg;
          double sum = 0.0, sq = 0.0;
          for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {
            const longlong2 st = __ldg(reinterpret_cast<const longlong2*>(
                (cc < p.gn_C0) ? p.gn_st0 + (static_cast<size_t>(b) * p.gn_C0 + cc) * 2
                               : p.gn_st1 + (static_cast<size_t>(b) * p.gn_C1 + (cc - p.gn_C0)) * 2));
            sum += static_cast<double>(st.x) * (1.0 / kStatSumScale);
            sq += static_cast<double>(st.y) * (1.0 / kStatSqScale);
          }
          const double mean = sum * inv_cnt;
          double var = sq * inv_cnt - mean * mean;
          if (var < 0.0) var = 0.0;
          const float rstd = rsqrtf(static_cast<float>(var) + p.gn_eps);
          const float sc = p.gn_gamma[c] * rstd;
          gn_s[c] = sc;
          gn_s[C::GN_MAXC + c] = p.gn_beta[c] - static_cast<float>(mean) * sc;
        }
        gn_b = b;
        named_bar_sync(2, XT);
      }
      for (int sg = 0; sg < p.nseg; ++sg) {
        const ConvSeg& S = p.seg[sg];
        if (S.raw == nullptr) { ai += S.nchunks; continue; }
        const bool inl = p.gn_st0 != nullptr;
        const float* aff = inl ? gn_s + S.aff_c0 + v * V : S.aff + static_cast<size_t>(b) * 2 * S.aff_C + S.aff_c0 + v * V;
        const int aff_row = inl ? C::GN_MAXC : S.aff_C;  // distance between the scale row and the shift row
        for (int kc = 0; kc < S.nchunks; ++kc, ++ai) {
          float sc[V], sh[V];
#pragma unroll
          for (int j = 0; j < V; ++j) sc[j] = sh[j] = 0.f;
          if (inside != 0) {  // (a ghost tile has no pixel inside the image and no sample to take the table from)
#pragma unroll
            for (int j = 0; j < V; j += 4) {
              float4 a, c;
              if (inl) {
                a = *reinterpret_cast<const float4*>(aff + kc * C::CK + j);
                c = *reinterpret_cast<const float4*>(aff + aff_row + kc * C::CK + j);
              } else {
                a = __ldg(reinterpret_cast<const float4*>(aff + kc * C::CK + j));
                c = __ldg(reinterpret_cast<const float4*>(aff + aff_row + kc * C::CK + j));
              }
              sc[j] = a.x; sc[j + 1] = a.y; sc[j + 2] = a.z; sc[j + 3] = a.w;
              sh[j] = c.x; sh[j + 1] = c.y; sh[j + 2] = c.z; sh[j + 3] = c.w;
            }
          }
          const uint32_t as = ai % C::A_SLOTS;
          mbar_wait_p<PROF>(&a_raw[as], (rawph >> as) & 1u, w_r);
          rawph ^= 1u << as;
          uint8_t* slot = sA + as * C::A_SLOT;  // 1024-byte aligned: the swizzle phase of window pixel q is q & 7
          bool dbg_noxf = false;
          if constexpr (PROF) dbg_noxf = (p.dbg & 4) != 0;
#pragma unroll
          for (int i = 0; i < NIT; ++i) {
            if (((inside >> i) & 1u) && !dbg_noxf) {
              const int q = pb + PSTEP * i;
              uint4* ptr = reinterpret_cast<uint4*>(slot + q * 128 + ((v ^ (q & 7)) << 4));
              float f[V];
              Vec<T>::unpack(*ptr, f);
#pragma unroll
              for (int j = 0; j < V; ++j) f[j] = silu_act<T>(fmaf(f[j], sc[j], sh[j]));
              *ptr = Vec<T>::pack_operand(f);
            }
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
          named_bar_sync(2, XT);
          if (tt == 0) {
            if constexpr (CG2) mbar_arrive_cluster(mapa_u32(smem_u32(&a_full[as]), 0));
            else mbar_arrive(&a_full[as]);
          }
        }
      }
    }
    if constexpr (PROF) {
      if (tt == 0) {
        atomicAdd(p.prof + 9, static_cast<unsigned long long>(w_r));
        atomicAdd(p.prof + 10, static_cast<unsigned long long>(clock64() - t_begin));
      }
    }
    }  // FUSE
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
    long long w_f = 0;
    const long long t_begin = PROF ? clock64() : 0;
    for (int unit = T0; unit < TEND; unit += TSTEP, ++ti) {
      const int tile = p.tile_base + unit / nsp;
      const int nb0 = (unit % nsp) * N;  // first output channel of this unit's slice (0 without an N-split)
      const int ldn = SWAP ? N : p.ldn;         // channel pitch of out / res / stats (the swap-AB form is never split)
      const bool ghost = tile >= p.ntiles;
      const int b = ghost ? 0 : tile / tiles_per_img;
      const int rem = tile - (tile / tiles_per_img) * tiles_per_img;
      const int th = rem / p.tiles_w;
      const int w = (rem - th * p.tiles_w) * C::TILE_W + wl;
      const int h = th * C::TILE_H + hl;
      const bool valid = !ghost && (h < p.H) && (w < p.W);
      const size_t pix = (static_cast<size_t>(b) * p.H + h) * p.W + w;
      const float* bias = p.bias + static_cast<size_t>(b) * p.bias_bstride + nb0;
      const uint32_t acs = ti & 1, acph = (ti >> 1) & 1;
      if (res != nullptr && !ghost) {
        // pull this thread's share of the residual tile into L2 while the MMAs of the tile are still running
        const int tw0p = (rem - th * p.tiles_w) * C::TILE_W, th0p = th * C::TILE_H;
        constexpr int ET = 32 * C::EPI_WARPS;
        constexpr int row_bytes = N * static_cast<int>(sizeof(T));  // one pixel
        constexpr int SEGS = (row_bytes + 127) / 128;
        for (int i = threadIdx.x - 64; i < C::TILE_H * C::TILE_W * SEGS; i += ET) {
          const int pixl = i / SEGS, seg = i % SEGS;
          const int hh = th0p + (pixl >> 3), ww = tw0p + (pixl & 7);
          if (hh < p.H && ww < p.W) {
            const char* a = reinterpret_cast<const char*>(res) +
                            (((static_cast<size_t>(b) * p.H + hh) * p.W + ww) * ldn + nb0) * sizeof(T) + seg * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          }
        }
      }
      // (the swap-AB forms request the first chunk of the residual BEFORE waiting for the accumulator)
      auto wait_acc = [&]() {
        mbar_wait_p<PROF>(&t_full[acs], acph, w_f);
        tc_fence_after();
      };
      bool dbg_skip = false, dbg_nostore = false;
      if constexpr (PROF) { dbg_skip = (p.dbg & 1) != 0; dbg_nostore = (p.dbg & 2) != 0; }
      if (dbg_skip) {
        wait_acc();
      } else if constexpr (SWAP) {
        // channel-major accumulator: this thread = output channel quad*32 + lane; warp half `sub` owns pixel columns
        // [128*sub, 128*sub + 128) of the 256-pixel tile = tile rows [16*sub, 16*sub + 16).  Every pixel is written
        // by one warp-wide store of 32 consecutive channels (64 B bf16 / 128 B fp32): no transposition needed, and the
        // GroupNorm statistics of a channel are a plain in-register sum over the thread's pixels.
        const int tw0 = (rem - th * p.tiles_w) * C::TILE_W;
        const int th0 = th * C::TILE_H;
        const int c = quad * 32 + lane;
        const float bias_c = __ldg(bias + c);
        const uint32_t tcol = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acs * C::ACC_COLS + sub * 128;
        const size_t rowstride = static_cast<size_t>(p.W) * N;
        const size_t tile_off = (static_cast<size_t>(b) * p.H + th0) * rowstride + static_cast<size_t>(tw0) * N + c;
        T* obase = out + tile_off;
        const T* rbase = res + tile_off;
        uint32_t wmask = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) wmask |= (tw0 + j < p.W) ? (1u << j) : 0u;
        if (ghost) wmask = 0;
        if constexpr (kBf16) {
          // bf16: lanes (2k, 2k+1) = channels (c, c+1) trade every other pixel so that each lane stores packed
          // bf16x2 words (128 B per warp-wide store instead of 64 B; measured: 2-byte stores cost 17 % of conv time)
          const bool odd = (lane & 1) != 0;
          const int ce = c & ~1;  // even channel of the pair
          __nv_bfloat16* ob2 = reinterpret_cast<__nv_bfloat16*>(out) + (tile_off - c + ce);
          const __nv_bfloat16* rb2 = reinterpret_cast<const __nv_bfloat16*>(res) + (tile_off - c + ce);
          float s2[2] = {0.f, 0.f}, q2[2] = {0.f, 0.f};  // partial sums for channels ce, ce+1 over this lane's pixels
          // residual words of the current 32-pixel chunk: chunk 0 is requested before the wait for the accumulator, the
          // others at the top of their chunk.  (The element-wise rolling prefetch of the fp32 form below was measured
          // here too: 16 % slower in bf16, where every element already costs a shuffle and a packed store.)
          uint32_t rres[16];
          auto load_res = [&](int ch, int k) -> uint32_t {
            const int row0 = sub * 16 + ch * 4;
            const int idx = 2 * k + (odd ? 1 : 0), i = idx >> 3, j = idx & 7;
            const bool ok = (th0 + row0 + i < p.H) && ((wmask >> j) & 1u);
            return ok ? __ldg(reinterpret_cast<const uint32_t*>(rb2 + (row0 + i) * rowstride + static_cast<size_t>(j) * N)) : 0u;
          };
          if (res != nullptr) {
#pragma unroll
            for (int k = 0; k < 16; ++k) rres[k] = load_res(0, k);
          }
          wait_acc();
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {
            const int row0 = sub * 16 + ch * 4;
            if (res != nullptr && ch > 0) {
#pragma unroll
              for (int k = 0; k < 16; ++k) rres[k] = load_res(ch, k);
            }
            uint32_t r[32];
            tmem_ld32(tcol + ch * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float mine_e = __uint_as_float(r[2 * k]) + bias_c, mine_o = __uint_as_float(r[2 * k + 1]) + bias_c;
              const float recv = __shfl_xor_sync(0xffffffffu, odd ? mine_e : mine_o, 1);
              // this lane's pixel: 2k (even lane) / 2k+1 (odd lane); lo = channel ce, hi = channel ce+1
              float lo = odd ? recv : mine_e, hi = odd ? mine_o : recv;
              if (res != nullptr) {
                lo += __uint_as_float(rres[k] << 16);
                hi += __uint_as_float(rres[k] & 0xffff0000u);
              }
              lo *= p.scale;
              hi *= p.scale;
              const int idx = 2 * k + (odd ? 1 : 0), i = idx >> 3, j = idx & 7;
              if ((th0 + row0 + i < p.H) && ((wmask >> j) & 1u)) {
                const __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
                if (!dbg_nostore) *reinterpret_cast<__nv_bfloat162*>(ob2 + (row0 + i) * rowstride + static_cast<size_t>(j) * N) = pk;
                s2[0] += lo; q2[0] = fmaf(lo, lo, q2[0]);
                s2[1] += hi; q2[1] = fmaf(hi, hi, q2[1]);
              }
            }
          }
          if (p.stats_acc != nullptr) {
            // even lane finishes channel ce (its own pixels + the partner's), odd lane channel ce+1; fixed order
            const float os = __shfl_xor_sync(0xffffffffu, odd ? s2[0] : s2[1], 1);
            const float oq = __shfl_xor_sync(0xffffffffu, odd ? q2[0] : q2[1], 1);
            const float ts = odd ? (os + s2[1]) : (s2[0] + os);  // (even-lane pixels) + (odd-lane pixels)
            const float tq = odd ? (oq + q2[1]) : (q2[0] + oq);
            stat_s[(sub * N + c) * 2] = ts;
            stat_s[(sub * N + c) * 2 + 1] = tq;
          }
        } else {
        float s_sum = 0.f, s_sq = 0.f;
        // residual of the current chunk, rolling prefetch: chunk 0 is requested before the wait for the accumulator and
        // element (i, j) of chunk ch + 1 as soon as element (i, j) of chunk ch has been consumed, so a chunk's loads are in
        // flight during the stores of the previous one (TF32 128 -> 128 + residual: 665 -> 713 TFLOP/s)
        float rv[32];
        auto load_res = [&](int ch, int i, int j) -> float {
          const int row0 = sub * 16 + ch * 4;
          const bool ok = (th0 + row0 + i < p.H) && ((wmask >> j) & 1u);
          return ok ? static_cast<float>(__ldg(rbase + (row0 + i) * rowstride + static_cast<size_t>(j) * N)) : 0.f;
        };
        if (res != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[i * 8 + j] = load_res(0, i, j);
        }
        wait_acc();
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          const int row0 = sub * 16 + ch * 4;  // first tile row of this 32-pixel chunk (4 rows x 8 columns)
          uint32_t r[32];
          tmem_ld32(tcol + ch * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool rok = th0 + row0 + i < p.H;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float v = __uint_as_float(r[i * 8 + j]) + bias_c;
              if (res != nullptr) {
                v += rv[i * 8 + j];
                if (ch < 3) rv[i * 8 + j] = load_res(ch + 1, i, j);
              }
              v *= p.scale;
              if (rok && ((wmask >> j) & 1u)) {
                obase[(row0 + i) * rowstride + static_cast<size_t>(j) * N] = static_cast<T>(v);
                s_sum += v;
                s_sq = fmaf(v, v, s_sq);
              }
            }
          }
        }
        if (p.stats_acc != nullptr) {
          // stat_s layout in this mode: [half = sub][N][2]
          stat_s[(sub * N + c) * 2] = s_sum;
          stat_s[(sub * N + c) * 2 + 1] = s_sq;
        }
        }  // fp32 direct stores
      } else {
      wait_acc();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acs * C::ACC_COLS + sub * N;
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 32) {
        constexpr int V = DT<T>::kVec;
        uint4 rq[32 / V];
        if (res != nullptr && valid) {
#pragma unroll
          for (int j = 0; j < 32 / V; ++j) rq[j] = __ldg(reinterpret_cast<const uint4*>(res + pix * ldn + nb0 + c0) + j);
        }
        uint32_t r[32];
        tmem_ld32(trow + c0, r);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
          f[j] = __uint_as_float(r[j]) + bb.x;
          f[j + 1] = __uint_as_float(r[j + 1]) + bb.y;
          f[j + 2] = __uint_as_float(r[j + 2]) + bb.z;
          f[j + 3] = __uint_as_float(r[j + 3]) + bb.w;
        }
        if (res != nullptr && valid) {
#pragma unroll
          for (int j = 0; j < 32; j += V) {
            float rr[V];
            Vec<T>::unpack(rq[j / V], rr);
#pragma unroll
            for (int q = 0; q < V; ++q) f[j + q] += rr[q];
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] *= p.scale;  // statistics below use these fp32 values (pre-rounding)
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += V) {
            float v[V];
#pragma unroll
            for (int q = 0; q < V; ++q) v[q] = f[j + q];
            Vec<T>::store(out + pix * ldn + nb0 + c0 + j, v);
          }
        }
        if (p.stats_acc != nullptr) {
          // column sums over this warp's 32 rows by recursive halving (31 shuffles): lane j ends with column c0 + j
          float a[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] = valid ? f[j] : 0.f;
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool hi = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < off; ++j) {
              const float send = hi ? a[j] : a[j + off];
              const float keep = hi ? a[j + off] : a[j];
              a[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          const float colsum = a[0];
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] = valid ? f[j] * f[j] : 0.f;
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool hi = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < off; ++j) {
              const float send = hi ? a[j] : a[j + off];
              const float keep = hi ? a[j + off] : a[j];
              a[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          stat_s[(ew * N + c0 + lane) * 2] = colsum;
          stat_s[(ew * N + c0 + lane) * 2 + 1] = a[0];
        }
      }
      }  // !SWAP
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG2) mbar_arrive_cluster(mapa_u32(smem_u32(&t_empty[acs]), 0));
        else mbar_arrive(&t_empty[acs]);
      }
      if (p.stats_acc != nullptr) {
        // combine the epilogue warps of this tile in a fixed order (deterministic) and publish the tile partial
        constexpr int ET = 32 * C::EPI_WARPS;
        asm volatile("bar.sync 1, %0;" ::"r"(ET) : "memory");
        long long* dst = p.stats_acc + (static_cast<size_t>(b) * ldn + nb0) * 2;
        for (int i = threadIdx.x - 64; i < N && !ghost; i += ET) {
          float sm_ = 0.f, sq_ = 0.f;
          if constexpr (SWAP) {
            sm_ = stat_s[i * 2] + stat_s[(N + i) * 2];  // the two pixel halves, fixed order
            sq_ = stat_s[i * 2 + 1] + stat_s[(N + i) * 2 + 1];
          } else {
#pragma unroll
            for (int w = 0; w < C::EPI_WARPS; ++w) { sm_ += stat_s[(w * N + i) * 2]; sq_ += stat_s[(w * N + i) * 2 + 1]; }
          }
          stat_atomic_add(dst + i * 2, sm_, sq_);  // integer atomics: order-independent, hence deterministic
        }
        asm volatile("bar.sync 1, %0;" ::"r"(ET) : "memory");
      }
    }
    if constexpr (PROF) {
      if (threadIdx.x == 64) {
        atomicAdd(p.prof + 7, static_cast<unsigned long long>(w_f));
        atomicAdd(p.prof + 8, static_cast<unsigned long long>(clock64() - t_begin));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (MC > 1) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
  if (warp == 1) {
    if constexpr (CG2) tmem_dealloc_cg2(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace use
