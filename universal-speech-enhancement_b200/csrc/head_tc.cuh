// Pyramid-head convolution: 3x3 pad 1, C -> pc (4 or 2) channels, fp32 output (+ FIR-upsample of the previous pyramid
// level), ncsnpp.py:440-461.  The operand is the already normalised + activated tensor (launch_gn_apply) or, FUSE, the
// RAW ResBlock output: six transform warps apply GroupNorm (per-sample scale / shift table) + SiLU + operand rounding in
// place in shared memory behind the TMA, exactly like the convolution's fused operand path (conv_tc.cuh) -- the
// normalised tensor of `GroupNorm -> SiLU -> conv3x3 C->4` never exists in HBM (one write + one read of C x H x W saved
// per level; bit-identical to gn_apply + the plain head).
//
// A C -> 4 convolution has no N dimension to speak of: as nine shifted MMAs it is bound by reading the activation tile
// from shared memory nine times (measured: 7.5 ms at 512 x 640 x 32 clips, 16 % of HBM speed).  Here the nine taps are
// folded into the N dimension instead:
//     D'[pixel][tap * pc + co] = sum_c A[pixel][c] * W[co][c][tap]          one MMA per K step, N = 48 (36 used)
//     out[y][x][co] = bias[co] + sum_{r,s} D'[(y + r - 1, x + s - 1)][(3 r + s) * pc + co]
// so every activation element is read from shared memory ONCE; the 3x3 gather happens on the 36 fp32 partial sums per
// pixel in the epilogue.  Tile = 16 x 16 window (TMA box, zero fill outside the image) -> 14 x 14 outputs.  The packed
// weights of all channel chunks stay resident in shared memory.  HBM-bound: one read of the operand.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer / TMEM owner, warps 2..17 = TWO epilogue groups of 8 warps
// (thread = window pixel).  A tile's epilogue is a serial chain (accumulator ready -> tcgen05.ld -> stage in shared memory
// -> barrier -> 3x3 gather -> store -> barrier, ~2.3 us) that one group could not hide behind the loads once the operand is
// bf16 (half the bytes per tile: round 1 measured 48 % of the HBM roofline in bf16 vs 86 % in fp32).  Group g owns
// accumulator slot g and staging buffer g and takes every other tile of the CTA, so two epilogues are in flight.
#pragma once
#include "common.cuh"

namespace use {

struct alignas(64) HeadParams {
  CUtensorMap tmA;  // rank 4 {C, W, H, B}, box {CK, 16, 16, 1}
  CUtensorMap tmW;  // rank 2 {C, 48}, box {CK, 48}
  int nchunks;      // C / CK  (<= kHeadMaxChunks)
  int B, H, W;
  int tiles_w, tiles_h, ntiles;
  const float* bias;   // [pc]
  float* out4;         // fp32 [B][H][W][pc]
  const float* prev4;  // optional fp32 [B][H/2][W/2][pc]
  int pc;
  const float* aff;    // FUSE: fp32 [B][2][C] scale row / shift row (launch_gn_affine)
  int C;               // FUSE: row length of aff
};

constexpr int kHeadMaxChunks = 8;
constexpr int kHeadN = 48;           // 9 taps x 4 outputs, padded to a multiple of 16
constexpr int kHeadTile = 14;        // outputs per tile edge
constexpr int kHeadWin = 16;         // window edge
constexpr int kHeadASlot = kHeadWin * kHeadWin * 128;  // 32 KB
constexpr int kHeadASlots = 3;
constexpr int kHeadWBytes = kHeadMaxChunks * kHeadN * 128;  // 48 KB
constexpr int kHeadStagePitch = 37;  // floats per pixel in the staging buffer (odd: conflict-free)
constexpr int kHeadStageBytes = 256 * kHeadStagePitch * 4;  // per epilogue group
constexpr int kHeadThreads = 64 + 2 * 256;
constexpr int kHeadXfThreads = 256;  // FUSE: eight transform warps behind the epilogue groups (six cannot keep up with a bf16 window per 1200 cycles)
constexpr int kHeadSmem = 1024 + kHeadASlots * kHeadASlot + kHeadWBytes + 2 * kHeadStageBytes + (3 * kHeadASlots + 5) * 8 + 16;
static_assert(kHeadSmem <= 232448, "shared memory budget");

template <typename T, bool FUSE>
__global__ void __launch_bounds__(kHeadThreads + (FUSE ? kHeadXfThreads : 0), 1) head_tc_kernel(const __grid_constant__ HeadParams p) {
  constexpr bool kBf16 = DT<T>::kIsBf16;
  constexpr int CK = 128 / sizeof(T);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW = sA + kHeadASlots * kHeadASlot;
  float* stage = reinterpret_cast<float*>(sW + kHeadWBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + kHeadWBytes + 2 * kHeadStageBytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kHeadASlots;
  uint64_t* a_raw = a_empty + kHeadASlots;  // FUSE: raw window landed (TMA -> transform warps)
  uint64_t* w_full = a_raw + kHeadASlots;
  uint64_t* t_full = w_full + 1;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmW);
    for (int i = 0; i < kHeadASlots; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_raw[i], 1); }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int g0 = blockIdx.x, gstep = gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, p.nchunks * kHeadN * 128);
      for (int kc = 0; kc < p.nchunks; ++kc) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(smem_u32(sW + kc * kHeadN * 128)), "l"(reinterpret_cast<uint64_t>(&p.tmW)), "r"(smem_u32(w_full)),
              "r"(kc * CK), "r"(0)
            : "memory");
      }
      uint32_t ai = 0;
      for (int tile = g0; tile < p.ntiles; tile += gstep) {
        const int b = tile / tiles_per_img;
        const int rem = tile - b * tiles_per_img;
        const int th = rem / p.tiles_w;
        const int w0 = (rem - th * p.tiles_w) * kHeadTile, h0 = th * kHeadTile;
        for (int kc = 0; kc < p.nchunks; ++kc, ++ai) {
          const uint32_t as = ai % kHeadASlots, aph = (ai / kHeadASlots) & 1;
          mbar_wait(&a_empty[as], aph ^ 1);
          uint64_t* landed = FUSE ? &a_raw[as] : &a_full[as];
          mbar_arrive_expect_tx(landed, kHeadASlot);
          tma_load_4d(sA + as * kHeadASlot, &p.tmA, landed, kc * CK, w0 - 1, h0 - 1, b);
        }
      }
    }
  } else if (warp == 1) {
    {  // the whole warp walks the loop; one elected lane issues (uniform control flow keeps the descriptors in uniform registers)
      constexpr uint32_t idesc = umma_idesc(kBf16 ? 1 : 2, 128, kHeadN);
      const uint32_t sA_addr = smem_u32(sA), sW_addr = smem_u32(sW);
      mbar_wait(w_full, 0);
      uint32_t ai = 0, ti = 0;
      for (int tile = g0; tile < p.ntiles; tile += gstep, ++ti) {
        const uint32_t acs = ti & 1, acph = (ti >> 1) & 1;
        mbar_wait(&t_empty[acs], acph ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < p.nchunks; ++kc, ++ai) {
          const uint32_t as = ai % kHeadASlots, aph = (ai / kHeadASlots) & 1;
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t ad = umma_desc_sw128(sA_addr + as * kHeadASlot + sub * 16384 + k * 32);
                const uint64_t bd = umma_desc_sw128(sW_addr + kc * kHeadN * 128 + k * 32);
                umma_ss<kBf16>(tmem_base + acs * 128 + sub * 64, ad, bd, idesc, (kc == 0 && k == 0) ? 0u : 1u);
              }
            }
            umma_commit(&a_empty[as]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&t_full[acs]);
        __syncwarp();
      }
    }
  } else if (threadIdx.x >= kHeadThreads) {
    if constexpr (FUSE) {
      // ================================ transform warps ================================
      // thread = one 16-byte channel vector (v) of the window pixels pb, pb + 32, ...; the window is 16 pixels x 128 B per
      // row = 1024-byte aligned rows of 8 pixels, so the swizzle phase of window pixel q is q & 7.  Pixels outside the
      // image keep the TMA's zero fill: the convolution pads the ACTIVATED tensor.
      constexpr int V = DT<T>::kVec;
      constexpr int PSTEP = kHeadXfThreads / 8;
      constexpr int NPIX = kHeadWin * kHeadWin;
      constexpr int NIT = (NPIX + PSTEP - 1) / PSTEP;
      const int tt = threadIdx.x - kHeadThreads;
      const int v = tt & 7, pb = tt >> 3;
      uint32_t ai = 0;
      for (int tile = g0; tile < p.ntiles; tile += gstep) {
        const int b = tile / tiles_per_img;
        const int rem = tile - b * tiles_per_img;
        const int th = rem / p.tiles_w;
        const int w0 = (rem - th * p.tiles_w) * kHeadTile, h0 = th * kHeadTile;
        uint32_t inside = 0;
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          const int q = pb + PSTEP * i;
          const int hh = h0 - 1 + (q >> 4), ww = w0 - 1 + (q & 15);
          if (q < NPIX && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) inside |= 1u << i;
        }
        const float* aff = p.aff + static_cast<size_t>(b) * 2 * p.C + v * V;
        for (int kc = 0; kc < p.nchunks; ++kc, ++ai) {
          float sc[V], sh[V];
#pragma unroll
          for (int j = 0; j < V; j += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(aff + kc * CK + j));
            const float4 c = __ldg(reinterpret_cast<const float4*>(aff + p.C + kc * CK + j));
            sc[j] = a.x; sc[j + 1] = a.y; sc[j + 2] = a.z; sc[j + 3] = a.w;
            sh[j] = c.x; sh[j + 1] = c.y; sh[j + 2] = c.z; sh[j + 3] = c.w;
          }
          const uint32_t as = ai % kHeadASlots, aph = (ai / kHeadASlots) & 1;
          mbar_wait(&a_raw[as], aph);
          uint8_t* slot = sA + as * kHeadASlot;
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
          fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
          named_bar_sync(3, kHeadXfThreads);
          if (tt == 0) mbar_arrive(&a_full[as]);
        }
      }
    }
  } else {
    const int grp = (warp - 2) >> 3;  // epilogue group: tiles ti = grp, grp + 2, ... of this CTA, accumulator slot grp
    const int ew = (warp - 2) & 7;
    const int sub = ew >> 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int pixel = sub * 128 + quad * 32 + lane;  // window pixel = accumulator row
    const int e = ew * 32 + lane;                    // gather role: output pixel e of the 14 x 14 tile (e < 196)
    const int oy = e / kHeadTile, ox = e - oy * kHeadTile;
    const int pc = p.pc;
    const int nd = 9 * pc;
    stage += grp * (256 * kHeadStagePitch);
    const uint32_t bar_id = 1 + grp;
    for (uint32_t ti = grp, tile = g0 + grp * gstep; tile < static_cast<uint32_t>(p.ntiles); tile += 2 * gstep, ti += 2) {
      const int b = tile / tiles_per_img;
      const int rem = tile - b * tiles_per_img;
      const int th = rem / p.tiles_w;
      const int h = th * kHeadTile + oy, w = (rem - th * p.tiles_w) * kHeadTile + ox;
      const uint32_t acs = ti & 1, acph = (ti >> 1) & 1;
      // FIR-upsampled previous pyramid level: request its four taps BEFORE waiting for the accumulator -- the epilogue of
      // a tile is serial (stage -> barrier -> gather -> barrier), a global-load latency inside it costs ~30 % of the tile
      const bool live = e < kHeadTile * kHeadTile && h < p.H && w < p.W;
      float up[4] = {0.f, 0.f, 0.f, 0.f};
      if (live && p.prev4 != nullptr) {
        const int Hp = p.H >> 1, Wp = p.W >> 1;
        const int my = h >> 1, mx = w >> 1;
        const int ya = (h & 1) ? my : my - 1, xa = (w & 1) ? mx : mx - 1;
        const float wya = (h & 1) ? 0.75f : 0.25f, wxa = (w & 1) ? 0.75f : 0.25f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const int yy = ya + dy;
          if (yy < 0 || yy >= Hp) continue;
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int xx = xa + dx;
            if (xx < 0 || xx >= Wp) continue;
            const float kw = (dy ? 1.f - wya : wya) * (dx ? 1.f - wxa : wxa);
            const float* pv = p.prev4 + ((static_cast<size_t>(b) * Hp + yy) * Wp + xx) * pc;
            if (pc == 4) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(pv));
              up[0] += kw * q.x; up[1] += kw * q.y; up[2] += kw * q.z; up[3] += kw * q.w;
            } else {
              const float2 q = __ldg(reinterpret_cast<const float2*>(pv));
              up[0] += kw * q.x; up[1] += kw * q.y;
            }
          }
        }
      }
      mbar_wait(&t_full[acs], acph);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acs * 128 + sub * 64;
      {
        uint32_t r[32];
        tmem_ld32(trow, r);
        tmem_ld_wait();
        float* dst = stage + pixel * kHeadStagePitch;
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(r[j]);
        if (nd > 32) {
          tmem_ld32(trow + 32, r);  // columns 32..63 of this 64-column slice (48 written by the MMA, 36 used)
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[32 + j] = __uint_as_float(r[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acs]);
      named_bar_sync(bar_id, 256);
      if (live) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int co = 0; co < 4; ++co) if (co < pc) o[co] = __ldg(p.bias + co);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const float* src = stage + ((oy + r) * kHeadWin + (ox + s)) * kHeadStagePitch + (r * 3 + s) * pc;
#pragma unroll
            for (int co = 0; co < 4; ++co) if (co < pc) o[co] += src[co];
          }
        }
#pragma unroll
        for (int co = 0; co < 4; ++co) o[co] += up[co];  // (conv + bias) + upsampled pyramid, like the reference's sum
        const size_t pix = (static_cast<size_t>(b) * p.H + h) * p.W + w;
        if (pc == 4) reinterpret_cast<float4*>(p.out4)[pix] = make_float4(o[0], o[1], o[2], o[3]);
        else reinterpret_cast<float2*>(p.out4)[pix] = make_float2(o[0], o[1]);
      }
      named_bar_sync(bar_id, 256);  // the staging buffer is rewritten by this group's next tile
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

}  // namespace use
