// Split-K form of the tcgen05 convolution for the LOW-RESOLUTION levels (<= 32 x 40 pixels per clip; 44 of the 99
// convolutions of an NCSN++ evaluation, layerspp.py:282-314 at ch_mult levels 4-6).
//
// Why: a tile of such a layer is one 128-pixel MMA row block whose whole K = 9 * C_in (+ skip) has to go through ONE SM.
// Measured (tools/conv_bench.py CONV_BENCH_SMALL, PROF counters, round 2): every tcgen05.mma of N <= 128 columns costs its
// issuer ~110 cycles whatever its width, so the K loop of conv_tc_kernel takes ~0.3 us per filter tap (36 - 72 taps =
// 11 - 22 us per launch) for 64-, 128- and 256-channel slices alike: slicing C_out shortens the epilogue, never the K loop.
// Here a CLUSTER of KS CTAs (2 or 4) shares one work unit (tile, 64-channel slice of C_out): CTA r accumulates taps
// [r K / KS, (r + 1) K / KS) of the linearised (segment, channel chunk, tap) sequence into its own TMEM accumulator;
// the partial sums P_1 .. P_{KS-1} are staged as fp32 in the owners' shared memory and the leader (rank 0) adds them to its
// own accumulator in rank order through distributed shared memory (ld.shared::cluster) before the usual epilogue
// (bias, residual, scale, rounding, GroupNorm statistics of the output).  The hand-off is two CLUSTER barriers per work
// unit, executed by every thread of every CTA of the cluster: #1 "all partial sums are staged" (arrive.release after a
// role's work on the unit, wait.acquire before the leader's remote reads), #2 "the leader has read them" (before a stage
// is rewritten or its CTA exits).  A first version signalled the leader with remote mbarrier arrives (release.cluster /
// try_wait.acquire.cluster) and kept the next unit's loads and MMAs running under the reduction; compute-sanitizer's
// racecheck does not model that as cluster-wide synchronisation (it flagged every staged word), and a latency-mode
// cluster almost always has ONE unit to do, so the barrier form costs nothing measurable and is checkable.
//
// Determinism / batch invariance: ((P0 + P1) + P2) + P3 is a fixed association, and WHICH layers run this form depends on
// the level geometry only (tiles per clip, conv_tc.cu), never on the batch: a clip sampled alone is still bit-identical to
// the same clip inside any batch.  Everything else -- window staging by TMA, the in-place GroupNorm + SiLU transform of a
// fused operand, descriptors, barriers -- is conv_tc_kernel's (conv_tc.cuh), specialised to NSUB = 1, pixel-major.
#pragma once
#include "conv_tc.cuh"

namespace use {

template <typename T, int N, bool FUSE, int KS>
struct ConvKsCfg {
  using Base = ConvCfg<T, N, 1, FUSE, false>;
  static constexpr int STAGE_BYTES = 128 * N * 4;  // this CTA's partial sums, fp32, [N / 4][128 pixels] float4 (conflict-free)
  static constexpr int KR_MAX = 64;                // channel chunks of one work unit (3 segments x <= 16 chunks)
  static constexpr int LIST_BYTES = KR_MAX * 16;
  static constexpr int NBARS = Base::NBARS;
  static constexpr int SMEM_BYTES = 1024 + Base::A_SLOTS * Base::A_SLOT + Base::B_SLOTS * Base::B_TILE + STAGE_BYTES +
                                    Base::STAT_BYTES + Base::GN_BYTES + LIST_BYTES + NBARS * 8 + 16;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(KS == 2 || KS == 4, "cluster width");
};

__device__ __forceinline__ float4 ld_cluster_f4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr));
  return v;
}

template <typename T, int N, bool FUSE, int KS>
__global__ void __launch_bounds__(ConvCfg<T, N, 1, FUSE, false>::THREADS, 1) conv_tc_ks_kernel(const __grid_constant__ ConvParams p) {
  using C = ConvCfg<T, N, 1, FUSE, false>;
  using K = ConvKsCfg<T, N, FUSE, KS>;
  constexpr bool kBf16 = DT<T>::kIsBf16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + C::A_SLOTS * C::A_SLOT;
  float4* stage = reinterpret_cast<float4*>(sB + C::B_SLOTS * C::B_TILE);
  float* stat_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stage) + K::STAGE_BYTES);
  float* gn_s = stat_s + C::STAT_BYTES / 4;  // [2][GN_MAXC] (FUSE)
  int4* klist = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(gn_s) + C::GN_BYTES);  // {segment, chunk, first tap, end tap}
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(klist) + K::LIST_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + C::A_SLOTS;
  uint64_t* a_raw = a_empty + C::A_SLOTS;
  uint64_t* b_full = a_raw + C::A_SLOTS;
  uint64_t* b_empty = b_full + C::B_SLOTS;
  uint64_t* t_full = b_empty + C::B_SLOTS;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  int* kr_n_s = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int krank = static_cast<int>(cluster_ctarank());

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nseg; ++i) {
      prefetch_tmap(&p.seg[i].tmA);
      prefetch_tmap(&p.seg[i].tmW);
    }
    for (int i = 0; i < C::A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_raw[i], 1); }
    for (int i = 0; i < C::B_SLOTS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], C::EPI_WARPS); }
    fence_barrier_init();
    // this CTA's share of the K sequence: taps [q0, q1) of the linearised (segment, chunk, tap) order
    int ktot = 0;
    for (int i = 0; i < p.nseg; ++i) ktot += p.seg[i].nchunks * p.seg[i].taps;
    const int q0 = krank * ktot / KS, q1 = (krank + 1) * ktot / KS;
    int n = 0, qb = 0;
    for (int sg = 0; sg < p.nseg; ++sg) {
      const int nt = p.seg[sg].taps;
      for (int kc = 0; kc < p.seg[sg].nchunks; ++kc, qb += nt) {
        const int t0 = max(q0 - qb, 0), t1 = min(q1 - qb, nt);
        if (t0 < t1 && n < K::KR_MAX) klist[n++] = make_int4(sg, kc, t0, t1);
      }
    }
    *kr_n_s = n;
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the whole cluster is resident before anything reads a peer's shared memory
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kr_n = *kr_n_s;
  pdl_wait();

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int T0 = blockIdx.x / KS, TSTEP = gridDim.x / KS;  // every CTA of a cluster walks the same work units
  const int TEND = p.nunits;
  const int nsp = p.nsplit;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // (lane 0 issues; the whole warp walks the unit loop because every thread takes part in the cluster barriers)
    int a_tile = T0, a_c = 0;
    uint32_t ai = 0;
    auto a_pending = [&]() { return a_tile < TEND; };
    auto a_issue = [&](bool blocking) -> bool {
      const uint32_t as = ai % C::A_SLOTS, aph = (ai / C::A_SLOTS) & 1;
      if (blocking) mbar_wait(&a_empty[as], aph ^ 1);
      else if (!mbar_test_wait(&a_empty[as], aph ^ 1)) return false;
      const int4 e = klist[a_c];
      const ConvSeg& S = p.seg[e.x];
      const int a_t = p.tile_base + a_tile / nsp;
      const int b = a_t / tiles_per_img;
      const int rem = a_t - b * tiles_per_img;
      const int th = rem / p.tiles_w;
      const int w0 = (rem - th * p.tiles_w) * C::TILE_W, h0 = th * C::TILE_H;
      const bool k3 = S.taps == 9;
      const uint32_t a_bytes = (k3 ? C::NPIX : C::TILE_H * 8) * 128;
      uint64_t* landed = (FUSE && S.raw != nullptr) ? &a_raw[as] : &a_full[as];
      mbar_arrive_expect_tx(landed, a_bytes);
      tma_load_4d(sA + as * C::A_SLOT, &S.tmA, landed, S.ac0 + e.y * C::CK, k3 ? (w0 - 1) : w0, k3 ? (h0 - 1) : h0, b);
      ++ai;
      if (++a_c == kr_n) { a_c = 0; a_tile += TSTEP; }
      return true;
    };
    uint32_t bi = 0, bj = 0;
    for (int unit = T0; unit < TEND; unit += TSTEP) {
      if (lane == 0) {
        const int wrow0 = (unit % nsp) * N;
        for (int c = 0; c < kr_n; ++c, ++bj) {
          const int4 e = klist[c];
          const ConvSeg& S = p.seg[e.x];
          const bool k3 = S.taps == 9;
          while (ai <= bj) a_issue(true);
          for (int tap = e.z; tap < e.w; ++tap) {
            if (a_pending() && ai < bj + C::A_SLOTS) a_issue(false);
            const uint32_t bs = bi % C::B_SLOTS, bph = (bi / C::B_SLOTS) & 1;
            mbar_wait(&b_empty[bs], bph ^ 1);
            const int wtap = k3 ? ((tap % 3) * 3 + tap / 3) : 0;  // tap order s-major, as in conv_tc_kernel
            mbar_arrive_expect_tx(&b_full[bs], C::B_TILE);
            tma_load_3d(sB + bs * C::B_TILE, &S.tmW, &b_full[bs], S.wc0 + e.y * C::CK, wrow0, wtap);
            ++bi;
          }
        }
      }
      __syncwarp();
      cluster_sync_all();  // #1: partial sums staged
      cluster_sync_all();  // #2: the leader has read them
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc = umma_idesc(kBf16 ? 1 : 2, 128, N);
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    uint32_t ai = 0, bi = 0, ti = 0;
    for (int unit = T0; unit < TEND; unit += TSTEP, ++ti) {
      const uint32_t acs = ti & 1, acph = (ti >> 1) & 1;
      mbar_wait(&t_empty[acs], acph ^ 1);
      tc_fence_after();
      bool first = true;
      for (int c = 0; c < kr_n; ++c, ++ai) {
        const int4 e = klist[c];
        const bool k3 = p.seg[e.x].taps == 9;
        const uint32_t sbo = k3 ? C::WIN_PITCH : 1024;
        const uint32_t as = ai % C::A_SLOTS, aph = (ai / C::A_SLOTS) & 1;
        mbar_wait(&a_full[as], aph);
        for (int tap = e.z; tap < e.w; ++tap, ++bi) {
          const int s = tap / 3, r = tap - s * 3;
          const uint32_t bs = bi % C::B_SLOTS, bph = (bi / C::B_SLOTS) & 1;
          mbar_wait(&b_full[bs], bph);
          tc_fence_after();
          const uint32_t win = sA_addr + as * C::A_SLOT + (k3 ? (r * C::WIN_W + s) * 128 : 0);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = umma_desc_sw128_sbo(win + k * 32, sbo);
              const uint64_t bd = umma_desc_sw128(sB_addr + bs * C::B_TILE + k * 32);
              umma_ss<kBf16>(tmem_base + acs * C::ACC_COLS, ad, bd, idesc, (first && k == 0) ? 0u : 1u);
            }
            umma_commit(&b_empty[bs]);
          }
          __syncwarp();
          first = false;
        }
        if (elect_one()) umma_commit(&a_empty[as]);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&t_full[acs]);
      __syncwarp();
      cluster_sync_all();  // #1
      cluster_sync_all();  // #2
    }
  } else if (threadIdx.x >= C::XF_T0) {
    if constexpr (FUSE) {
      // ================================ transform warps (conv_tc.cuh) ================================
      constexpr int V = DT<T>::kVec;
      constexpr int XT = C::XF_THREADS;
      constexpr int PSTEP = XT / 8;
      constexpr int NIT = (C::NPIX + PSTEP - 1) / PSTEP;
      const int tt = threadIdx.x - C::XF_T0;
      const int v = tt & 7, pb = tt >> 3;
      uint32_t ai = 0, rawph = 0;
      int gn_b = -1;
      for (int unit = T0; unit < TEND; unit += TSTEP) {
        const int tile = p.tile_base + unit / nsp;
        const int b = tile / tiles_per_img;
        const int rem = tile - b * tiles_per_img;
        const int th = rem / p.tiles_w;
        const int w0 = (rem - th * p.tiles_w) * C::TILE_W;
        const int h0 = th * C::TILE_H;
        uint32_t inside = 0;
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          const int q = pb + PSTEP * i;
          const int row = q / C::WIN_W, col = q - row * C::WIN_W;
          const int hh = h0 - 1 + row, ww = w0 - 1 + col;
          if (q < C::NPIX && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) inside |= 1u << i;
        }
        if (p.gn_st0 != nullptr && b != gn_b) {
          // inline GroupNorm table of a new sample: the arithmetic of gn_affine_kernel, bit for bit (conv_tc.cuh)
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
        const bool inl = p.gn_st0 != nullptr;
        for (int c = 0; c < kr_n; ++c, ++ai) {
          const int4 e = klist[c];
          const ConvSeg& S = p.seg[e.x];
          if (S.raw == nullptr) continue;
          const float* aff = inl ? gn_s + S.aff_c0 + v * V : S.aff + static_cast<size_t>(b) * 2 * S.aff_C + S.aff_c0 + v * V;
          const int aff_row = inl ? C::GN_MAXC : S.aff_C;
          const int kc = e.y;
          float sc[V], sh[V];
#pragma unroll
          for (int j = 0; j < V; j += 4) {
            float4 a, s4;
            if (inl) {
              a = *reinterpret_cast<const float4*>(aff + kc * C::CK + j);
              s4 = *reinterpret_cast<const float4*>(aff + aff_row + kc * C::CK + j);
            } else {
              a = __ldg(reinterpret_cast<const float4*>(aff + kc * C::CK + j));
              s4 = __ldg(reinterpret_cast<const float4*>(aff + aff_row + kc * C::CK + j));
            }
            sc[j] = a.x; sc[j + 1] = a.y; sc[j + 2] = a.z; sc[j + 3] = a.w;
            sh[j] = s4.x; sh[j + 1] = s4.y; sh[j + 2] = s4.z; sh[j + 3] = s4.w;
          }
          const uint32_t as = ai % C::A_SLOTS;
          mbar_wait(&a_raw[as], (rawph >> as) & 1u);
          rawph ^= 1u << as;
          uint8_t* slot = sA + as * C::A_SLOT;
#pragma unroll
          for (int i = 0; i < NIT; ++i) {
            if ((inside >> i) & 1u) {
              const int q = pb + PSTEP * i;
              uint4* ptr = reinterpret_cast<uint4*>(slot + q * 128 + ((v ^ (q & 7)) << 4));
              float f[V];
              Vec<T>::unpack(*ptr, f);
#pragma unroll
              for (int j = 0; j < V; ++j) f[j] = silu_act<T>(fmaf(f[j], sc[j], sh[j]));
              *ptr = Vec<T>::pack_operand(f);
            }
          }
          fence_proxy_async();
          named_bar_sync(2, XT);
          if (tt == 0) mbar_arrive(&a_full[as]);
        }
        cluster_sync_all();  // #1
        cluster_sync_all();  // #2
      }
    }
  } else {
    // ================================ epilogue ================================
    const int ew = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int m = quad * 32 + lane;
    const int hl = m >> 3, wl = m & 7;
    T* out = reinterpret_cast<T*>(p.out);
    const T* res = reinterpret_cast<const T*>(p.res);
    const int ldn = p.ldn;
    uint32_t ti = 0;
    uint32_t stage_rank[KS];  // shared::cluster addresses of the stage of every rank (rank 0 unused)
#pragma unroll
    for (int r = 0; r < KS; ++r) stage_rank[r] = mapa_u32(smem_u32(stage), r);
    for (int unit = T0; unit < TEND; unit += TSTEP, ++ti) {
      const uint32_t acs = ti & 1, acph = (ti >> 1) & 1;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acs * C::ACC_COLS;
      if (krank != 0) {
        // ---- owner of a partial sum: TMEM -> own shared memory ----
        mbar_wait(&t_full[acs], acph);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(trow + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            stage[((c0 + j) >> 2) * 128 + m] = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                           __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[acs]);
        __syncwarp();
        cluster_sync_all();  // #1: staged (release) -- the leader reads after its wait (acquire)
        cluster_sync_all();  // #2: the leader is done with this stage
        continue;
      }
      // ---- leader ----
      const int tile = p.tile_base + unit / nsp;
      const int nb0 = (unit % nsp) * N;
      const int b = tile / tiles_per_img;
      const int rem = tile - b * tiles_per_img;
      const int th = rem / p.tiles_w;
      const int w = (rem - th * p.tiles_w) * C::TILE_W + wl;
      const int h = th * C::TILE_H + hl;
      const bool valid = (h < p.H) && (w < p.W);
      const size_t pix = (static_cast<size_t>(b) * p.H + h) * p.W + w;
      const float* bias = p.bias + static_cast<size_t>(b) * p.bias_bstride + nb0;
      // pull the unit's bias row (and below: nothing else is read from global memory before the stores) towards the SM
      // while the MMAs still run: a single-unit launch would otherwise meet an L2 round trip per 32-column chunk
      if (lane < N / 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(bias + lane * 4));
      mbar_wait(&t_full[acs], acph);
      tc_fence_after();
      cluster_sync_all();  // #1: every owner's partial sums are staged
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
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]);
        // ((P0 + P1) + P2) + P3: fixed order
#pragma unroll
        for (int rk = 1; rk < KS; ++rk) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 q = ld_cluster_f4(stage_rank[rk] + static_cast<uint32_t>((((c0 + j) >> 2) * 128 + m) * 16));
            f[j] += q.x; f[j + 1] += q.y; f[j + 2] += q.z; f[j + 3] += q.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
          f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w;
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
        for (int j = 0; j < 32; ++j) f[j] *= p.scale;
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += V) {
            float vv[V];
#pragma unroll
            for (int q = 0; q < V; ++q) vv[q] = f[j + q];
            Vec<T>::store(out + pix * ldn + nb0 + c0 + j, vv);
          }
        }
        if (p.stats_acc != nullptr) {
          // column sums over this warp's 32 rows by recursive halving: lane j ends with column c0 + j (conv_tc.cuh)
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acs]);
      __syncwarp();
      cluster_arrive();  // #2 (arrive): the remote reads of this unit are done; the stores below need no peer
      if (p.stats_acc != nullptr) {
        constexpr int ET = 32 * C::EPI_WARPS;
        asm volatile("bar.sync 1, %0;" ::"r"(ET) : "memory");
        long long* dst = p.stats_acc + (static_cast<size_t>(b) * ldn + nb0) * 2;
        for (int i = threadIdx.x - 64; i < N; i += ET) {
          float sm_ = 0.f, sq_ = 0.f;
#pragma unroll
          for (int wq = 0; wq < C::EPI_WARPS; ++wq) { sm_ += stat_s[(wq * N + i) * 2]; sq_ += stat_s[(wq * N + i) * 2 + 1]; }
          stat_atomic_add(dst + i * 2, sm_, sq_);
        }
        asm volatile("bar.sync 1, %0;" ::"r"(ET) : "memory");
      }
      cluster_wait();  // #2 (wait)
    }
  }

  // (barrier #2 of the last unit is the exit guard: no owner leaves before the leader has read its stage)
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace use
