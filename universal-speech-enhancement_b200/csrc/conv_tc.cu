// Host side of the tcgen05 convolution: TMA tensor-map construction and launch.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "conv_tc.cuh"
#include "conv_tc_ks.cuh"
#include "head_tc.cuh"
#include "kernels.h"

namespace use {

// cuTensorMapEncodeTiled is fetched through the runtime so the library has no link-time dependency on
// libcuda.so (the CPU build box has none).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TcConvPlan {
  ConvParams params;
  int dt, N, nsub, mc;  // mc: cluster width of the launch (weight multicast pairs, CTA-pair MMAs or split-K clusters)
  int ks = 1;           // split-K cluster width (conv_tc_ks.cuh); 1: conv_tc_kernel
  int grid, threads, smem;
  const void* kernel;
  TcConvPlan* tail = nullptr;  // sliced launch over the last (ntiles mod SMs) tiles, run right after this one
};

static bool encode_act(CUtensorMap* m, int dt, const void* base, int B, int H, int W, int Ct, int cols, int rows,
                       char* err, int errlen) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled unavailable"); return false; }
  const cuuint64_t es = act_size(dt);
  const cuuint32_t ck = 128 / es;
  cuuint64_t dims[4] = {(cuuint64_t)Ct, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {Ct * es, (cuuint64_t)W * Ct * es, (cuuint64_t)H * W * Ct * es};
  cuuint32_t box[4] = {ck, (cuuint32_t)cols, (cuuint32_t)rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, dt == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled(act B=%d H=%d W=%d C=%d rows=%d) failed: %d", B, H, W, Ct, rows, (int)r);
    return false;
  }
  return true;
}

static bool encode_w(CUtensorMap* m, int dt, const void* base, int taps, int N, int Ctot, int box_rows, char* err,
                     int errlen) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled unavailable"); return false; }
  const cuuint64_t es = act_size(dt);
  const cuuint32_t ck = 128 / es;
  cuuint64_t dims[3] = {(cuuint64_t)Ctot, (cuuint64_t)N, (cuuint64_t)taps};
  cuuint64_t strides[2] = {Ctot * es, (cuuint64_t)N * Ctot * es};
  cuuint32_t box[3] = {ck, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, dt == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled(weights taps=%d N=%d C=%d) failed: %d", taps, N, Ctot, (int)r);
    return false;
  }
  return true;
}

// CTA-pair weight multicast (cluster of 2) is OFF by default: measured on B200 it changes nothing (bf16 128 -> 128 at
// 512 x 640 x 16: 1442 TFLOP/s with it, 1462 without).  The L2 -> SM path is the limiter of these layers (the PROF build
// gains 18 % with the weight stream removed), but a multicast to fewer than ~8 CTAs is not deduplicated in L2 on this
// part, so a pair saves no L2 bandwidth.  USE_B200_CONV_MC=2 enables it for A/B measurements.
static int multicast_width() {
  static int mc = [] {
    const char* v = getenv("USE_B200_CONV_MC");
    return (v && v[0] == '2') ? 2 : 1;
  }();
  return mc;
}

template <typename T, int N, int NSUB, bool SWAP, bool FUSE, int MC, bool PROF, bool CG2 = false>
static void set_kernel(TcConvPlan* p) {
  using C = ConvCfg<T, N, NSUB, FUSE, CG2>;
  auto k = &conv_tc_kernel<T, N, NSUB, SWAP, FUSE, MC, PROF, CG2>;
  p->kernel = reinterpret_cast<const void*>(k);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  p->threads = C::THREADS;
  p->smem = C::SMEM_BYTES;
  p->nsub = NSUB;
  p->mc = MC;
}

template <typename T, bool FUSE, int KS>
static void set_ks_kernel(TcConvPlan* p) {
  auto k = &conv_tc_ks_kernel<T, 64, FUSE, KS>;
  p->kernel = reinterpret_cast<const void*>(k);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvKsCfg<T, 64, FUSE, KS>::SMEM_BYTES);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
  p->threads = ConvCfg<T, 64, 1, FUSE, false>::THREADS;
  p->smem = ConvKsCfg<T, 64, FUSE, KS>::SMEM_BYTES;
  p->nsub = 1;
  p->mc = KS;
  p->ks = KS;
}
template <typename T>
static void fill_ks_kernel(TcConvPlan* p, bool fuse, int ks) {
  if (ks == 4) { if (fuse) set_ks_kernel<T, true, 4>(p); else set_ks_kernel<T, false, 4>(p); }
  else { if (fuse) set_ks_kernel<T, true, 2>(p); else set_ks_kernel<T, false, 2>(p); }
}

// CTA-pair MMAs (tcgen05 cta_group::2) for the C_out = 128 layers: USE_B200_CONV_CG2=1 / 0
static bool cg2_enabled() {
  static bool on = [] {
    const char* v = getenv("USE_B200_CONV_CG2");
    return v && v[0] == '1';
  }();
  return on;
}

template <typename T, int N, int NSUB>
static void fill_kernel(TcConvPlan* p, bool fuse) {
  constexpr bool SWAP = (N == 128 && NSUB == 2);  // swap-AB for C_out = 128 (conv_tc.cuh)
  const bool mc2 = multicast_width() == 2;
  if constexpr (N == 128 && NSUB == 2) {
    if (cg2_enabled()) {
      if (getenv("USE_B200_CONV_PROF")) {
        if (fuse) set_kernel<T, N, NSUB, false, true, 2, true, true>(p);
        else set_kernel<T, N, NSUB, false, false, 2, true, true>(p);
        if (const char* v = getenv("USE_B200_CONV_DBG")) p->params.dbg = atoi(v);
        cudaMalloc(&p->params.prof, 16 * sizeof(unsigned long long));
        cudaMemset(p->params.prof, 0, 16 * sizeof(unsigned long long));
        return;
      }
      if (fuse) set_kernel<T, N, NSUB, false, true, 2, false, true>(p);
      else set_kernel<T, N, NSUB, false, false, 2, false, true>(p);
      return;
    }
  }
  if constexpr (SWAP) {
    if (getenv("USE_B200_CONV_PROF")) {
      // instrumented build of the C_out = 128 kernel (tools/conv_bench.py): per-role mbarrier stall cycles
      if (fuse) { if (mc2) set_kernel<T, N, NSUB, SWAP, true, 2, true>(p); else set_kernel<T, N, NSUB, SWAP, true, 1, true>(p); }
      else { if (mc2) set_kernel<T, N, NSUB, SWAP, false, 2, true>(p); else set_kernel<T, N, NSUB, SWAP, false, 1, true>(p); }
      if (const char* v = getenv("USE_B200_CONV_DBG")) p->params.dbg = atoi(v);
      cudaMalloc(&p->params.prof, 16 * sizeof(unsigned long long));
      cudaMemset(p->params.prof, 0, 16 * sizeof(unsigned long long));
      return;
    }
  }
  if (fuse) { if (mc2) set_kernel<T, N, NSUB, SWAP, true, 2, false>(p); else set_kernel<T, N, NSUB, SWAP, true, 1, false>(p); }
  else { if (mc2) set_kernel<T, N, NSUB, SWAP, false, 2, false>(p); else set_kernel<T, N, NSUB, SWAP, false, 1, false>(p); }
}

bool tc_conv_supported(int dt, int N) { return N == 64 || N == 128 || N == 256; }

// tile_base / tile_count: the tile range of this launch (count < 0: all); force_split: run the range as 64-channel slices
static TcConvPlan* plan_create_range(int dt, const TcConvDesc& d, int num_sms, char* err, int errlen, int tile_base,
                                     int tile_count, bool force_split);

TcConvPlan* tc_conv_plan_create(int dt, const TcConvDesc& d, int num_sms, char* err, int errlen) {
  // Tail split (C_out = 256 family, single-CTA form): with T tiles on S persistent CTAs the last wave holds only T mod S
  // tiles (batch 1 at 128 x 160: 160 tiles on 148 SMs = two waves for 1.08 waves of work).  When that remainder is small,
  // the first T - (T mod S) tiles run as an exactly balanced launch and the remainder as 64-channel slices (4x the work
  // units, a quarter of the time each) right behind it.  Slices are bit-identical to whole tiles (see N-split).
  // MEASURED (bf16, batch 1 / 2 / 4): 159.9 vs 155.9, 222.4 vs 220.9, 393.3 vs 384.7 ms per step -- SLOWER: a second launch
  // costs ~10 us of fixed latency (TMEM allocation, barrier set-up, pipeline fill, drain), as much as the tile it saves.
  // Off unless USE_B200_CONV_TAILSPLIT=1.
  if (d.N == 256 && multicast_width() == 1 && !cg2_enabled() && !getenv("USE_B200_CONV_PROF")) {
    static const bool off = !(getenv("USE_B200_CONV_TAILSPLIT") && getenv("USE_B200_CONV_TAILSPLIT")[0] == '1');
    const int tiles = ((d.W + 7) / 8) * ((d.H + 15) / 16) * d.B;
    const int tail = tiles % num_sms;
    if (!off && tiles > num_sms && tail > 0 && tail * 3 <= num_sms) {
      TcConvPlan* main = plan_create_range(dt, d, num_sms, err, errlen, 0, tiles - tail, false);
      if (!main) return nullptr;
      main->tail = plan_create_range(dt, d, num_sms, err, errlen, tiles - tail, tail, true);
      if (!main->tail) { delete main; return nullptr; }
      return main;
    }
  }
  return plan_create_range(dt, d, num_sms, err, errlen, 0, -1, false);
}

static TcConvPlan* plan_create_range(int dt, const TcConvDesc& d, int num_sms, char* err, int errlen, int tile_base,
                                     int tile_count, bool force_split) {
  if (!tc_conv_supported(dt, d.N)) {
    snprintf(err, errlen, "tcgen05 conv: unsupported C_out=%d", d.N);
    return nullptr;
  }
  TcConvPlan* p = new TcConvPlan();
  memset(&p->params, 0, sizeof(p->params));
  p->dt = dt;
  p->N = d.N;
  bool fuse = false;
  for (int i = 0; i < d.nseg; ++i) fuse = fuse || d.seg[i].aff != nullptr;
  // N-split (conv_tc.cuh, ConvParams::nsplit): a launch with few tiles runs 64- or 128-channel slices of C_out as extra
  // work units, as many as still fit ONE wave of CTAs (first version, slices whenever tiles <= SMs / 3: batch 1
  // 162.2 -> 157.7 ms per clip; at 80-147 tiles the 4x activation re-reads cost as much as the shorter weight stream
  // saves: batch 2 unchanged, batch 4 2 % slower).
  // Only the C_out = 256 family: its unsplit kernel is the same pixel-major form, so split and unsplit launches are
  // bit-identical (a clip sampled alone == the same clip inside any batch).  The C_out = 128 layers run the swap-AB kernel,
  // whose channel-major epilogue sums the GroupNorm statistics in a different order: slicing them would make the result
  // depend on the batch size (measured: 1 ulp in the statistics, amplified to the rounding-noise floor).
  int nsplit = 1;
  {
    const int base_h = d.N == 256 ? 16 : 32;
    const int base_tiles = ((d.W + 7) / 8) * ((d.H + base_h - 1) / base_h) * d.B;
    static const bool off = getenv("USE_B200_CONV_NSPLIT") && getenv("USE_B200_CONV_NSPLIT")[0] == '0';
    static const int max_tiles = getenv("USE_B200_CONV_NSPLIT_MAXTILES") ? atoi(getenv("USE_B200_CONV_NSPLIT_MAXTILES")) : (1 << 30);
    // Slices are only worth it while ALL work units fit one wave: measured in isolation (bf16, batch 1, tools/conv_bench.py
    // CONV_BENCH_SMALL): in a single-tile K loop every tcgen05.mma of N <= 128 costs its issuer ~110 cycles whatever its
    // width (~0.3 us per tap for N = 64, 128 and 256 alike; alternating two accumulators changed nothing, so it is not the
    // accumulator dependency), so a unit costs the same whatever its width and a second wave doubles the launch.  64 x 80 (40
    // tiles): unsplit 21 us, four slices (160 units = 2 waves) 26 us -> two 128-channel slices (80 units, one wave).
    if (!off && d.N == 256 && base_tiles < max_tiles && multicast_width() == 1 && !cg2_enabled() && !getenv("USE_B200_CONV_PROF")) {
      if (base_tiles * 4 <= num_sms) nsplit = 4;
      else if (base_tiles * 2 <= num_sms) nsplit = 2;
    }
    // (Also slicing launches whose last wave is mostly empty -- batch 1 at 128 x 160: 160 tiles on 148 CTAs -- was measured:
    // batch 1 148.8 -> 162.4 ms, batch 2 214.1 -> 249.9 ms.  A slice repeats the window load AND the GroupNorm transform of
    // its tile, so four slices cost far more than the half-empty wave they fill.)
    if (force_split) nsplit = d.N / 64;
  }
  // Split-K clusters (conv_tc_ks.cuh) for the low-resolution levels of a LATENCY-mode program (TcConvDesc::latency, set
  // by the engine for jobs of at most two clips): <= 3 tiles per clip (16 x 20, 8 x 10): 4 CTAs per 64-channel work unit;
  // <= 10 (32 x 40): 2, so that one clip still fits one wave (10 tiles x 4 slices x 2 = 80 CTAs).  The partial-sum
  // association differs from conv_tc_kernel's single accumulator, hence a mode of the whole program and never a per-launch
  // heuristic: inside a mode, results do not depend on the batch.  Measured (bf16 / TF32, round 2): batch 1 140.0 -> 131.1 /
  // 235.5 -> 214.0 ms per clip, batch 2 215.9 -> 211.9 ms per step; but 16 clips 1394 -> 1421 ms and 32 clips (TF32)
  // 4990 -> 5103 ms per step (the 64-channel slices repeat window loads and GroupNorm transforms four times, the
  // reduction adds a cluster round trip per unit): throughput-mode programs keep the single-accumulator kernels.
  int ks = 1;
  {
    static const bool ks_off = getenv("USE_B200_CONV_KSPLIT") && getenv("USE_B200_CONV_KSPLIT")[0] == '0';
    const int tiles_img = ((d.W + 7) / 8) * ((d.H + 15) / 16);
    int ktot = 0;
    for (int i = 0; i < d.nseg; ++i) ktot += d.seg[i].taps * (d.seg[i].C / (128 / (int)act_size(dt)));
    static const int ks_max_tiles = getenv("USE_B200_CONV_KSPLIT_MAXTILES") ? atoi(getenv("USE_B200_CONV_KSPLIT_MAXTILES")) : 10;
    if (d.latency && !ks_off && !force_split && tile_count < 0 && d.N == 256 && tiles_img <= ks_max_tiles && ktot >= 4 && multicast_width() == 1 &&
        !cg2_enabled() && !getenv("USE_B200_CONV_PROF")) {
      ks = tiles_img <= 3 ? 4 : 2;
      nsplit = 4;
    }
  }
  if (getenv("USE_B200_CONV_DEBUG")) {
    const int bh = d.N == 256 ? 16 : 32;
    fprintf(stderr, "conv plan B=%d H=%d W=%d N=%d nseg=%d base_tiles=%d nsplit=%d ks=%d fuse=%d\n", d.B, d.H, d.W, d.N, d.nseg,
            ((d.W + 7) / 8) * ((d.H + bh - 1) / bh) * d.B, nsplit, ks, (int)fuse);
  }
  const int kN = d.N / nsplit;  // the kernel's N (MMA width, TMEM columns, weight-tile rows)
  if (ks > 1) {
    if (dt == kBF16) fill_ks_kernel<__nv_bfloat16>(p, fuse, ks);
    else fill_ks_kernel<float>(p, fuse, ks);
  } else if (dt == kBF16) {
    if (nsplit == 4) fill_kernel<__nv_bfloat16, 64, 1>(p, fuse);
    else if (nsplit == 2) fill_kernel<__nv_bfloat16, 128, 1>(p, fuse);
    else if (d.N == 256) fill_kernel<__nv_bfloat16, 256, 1>(p, fuse);
    else if (d.N == 128) fill_kernel<__nv_bfloat16, 128, 2>(p, fuse);
    else fill_kernel<__nv_bfloat16, 64, 2>(p, fuse);
  } else {
    if (nsplit == 4) fill_kernel<float, 64, 1>(p, fuse);
    else if (nsplit == 2) fill_kernel<float, 128, 1>(p, fuse);
    else if (d.N == 256) fill_kernel<float, 256, 1>(p, fuse);
    else if (d.N == 128) fill_kernel<float, 128, 2>(p, fuse);
    else fill_kernel<float, 64, 2>(p, fuse);
  }
  const int ck = 128 / (int)act_size(dt);
  const int tile_h = 16 * p->nsub;
  ConvParams& P = p->params;
  P.nseg = d.nseg;
  for (int i = 0; i < d.nseg; ++i) {
    const TcSegDesc& s = d.seg[i];
    if (s.C % ck || s.c0 % ck || s.wc0 % ck || (s.taps != 9 && s.taps != 1)) {
      snprintf(err, errlen, "tcgen05 conv: segment %d channels (C=%d c0=%d wc0=%d) not a multiple of %d or bad taps %d", i,
               s.C, s.c0, s.wc0, ck, s.taps);
      delete p;
      return nullptr;
    }
    if (s.aff != nullptr && s.taps != 9) {
      snprintf(err, errlen, "tcgen05 conv: only 3x3 segments can take a fused GroupNorm operand");
      delete p;
      return nullptr;
    }
    const int rows = s.taps == 9 ? tile_h + 2 : tile_h, cols = s.taps == 9 ? 10 : 8;
    if (!encode_act(&P.seg[i].tmA, dt, s.act, d.B, d.H, d.W, s.C_tensor, cols, rows, err, errlen) ||
        !encode_w(&P.seg[i].tmW, dt, s.w, s.taps, d.N, s.Cw_total, kN, err, errlen) ||
        !encode_w(&P.seg[i].tmWh, dt, s.w, s.taps, d.N, s.Cw_total, kN / 2, err, errlen)) {
      delete p;
      return nullptr;
    }
    P.seg[i].raw = s.aff != nullptr ? s.act : nullptr;
    P.seg[i].aff = s.aff;
    P.seg[i].Ct = s.C_tensor;
    P.seg[i].aff_C = s.aff_C;
    P.seg[i].aff_c0 = s.aff_c0;
    P.seg[i].nchunks = s.C / ck;
    P.seg[i].taps = s.taps;
    P.seg[i].wc0 = s.wc0;
    P.seg[i].ac0 = s.c0;
  }
  P.B = d.B; P.H = d.H; P.W = d.W;
  P.tiles_w = (d.W + 7) / 8;
  P.tiles_h = (d.H + tile_h - 1) / tile_h;
  P.ntiles = P.tiles_w * P.tiles_h * d.B;
  P.tile_base = tile_base;
  if (tile_count >= 0) P.ntiles = tile_base + tile_count;  // the END of this launch's range
  P.gn_st0 = fuse ? d.gn_st0 : nullptr;
  P.gn_st1 = d.gn_st1; P.gn_gamma = d.gn_gamma; P.gn_beta = d.gn_beta;
  P.gn_C0 = d.gn_C0; P.gn_C1 = d.gn_st1 ? d.gn_C1 : 0; P.gn_HW = d.gn_HW; P.gn_eps = d.gn_eps;
  if (P.gn_st0 && P.gn_C0 + P.gn_C1 > 512) {
    snprintf(err, errlen, "tcgen05 conv: inline GroupNorm over %d channels (max 512)", P.gn_C0 + P.gn_C1);
    delete p;
    return nullptr;
  }
  P.nsplit = nsplit;
  P.ldn = d.N;
  P.nunits = (P.ntiles - P.tile_base) * nsplit;
  P.out = d.out; P.bias = d.bias; P.bias_bstride = d.bias_bstride; P.res = d.res; P.scale = d.scale;
  P.stats_acc = d.stats_acc;
  if (p->ks > 1) {
    // one cluster per work unit, as many clusters as can be resident (one CTA per SM; GPC boundaries cost a few)
    int max_clusters = num_sms / p->ks;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(num_sms / p->ks * p->ks);
    cfg.blockDim = dim3(p->threads);
    cfg.dynamicSmemBytes = p->smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p->ks;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, p->kernel, &cfg) == cudaSuccess && n > 0 && n < max_clusters) max_clusters = n;
    cudaGetLastError();
    p->grid = (P.nunits < max_clusters ? P.nunits : max_clusters) * p->ks;
    return p;
  }
  const int units = (P.nunits + p->mc - 1) / p->mc;  // work-unit groups
  const int max_groups = num_sms / p->mc;
  p->grid = (units < max_groups ? units : max_groups) * p->mc;  // persistent CTAs, one per SM, a multiple of the cluster
  return p;
}

void tc_conv_plan_destroy(TcConvPlan* p) {
  if (p && p->tail) tc_conv_plan_destroy(p->tail);
  if (p && p->params.prof) {
    unsigned long long h[16];
    cudaMemcpy(h, p->params.prof, sizeof(h), cudaMemcpyDeviceToHost);
    const double g = p->grid;
    fprintf(stderr,
            "USE_B200_CONV_PROF per-CTA mean cycles: mma total %.0f wait[t_empty %.0f a_full %.0f b_full %.0f] | tma total %.0f "
            "wait[a_empty %.0f b_empty %.0f] | epi total %.0f wait[t_full %.0f] | xf total %.0f wait[a_raw %.0f]\n",
            h[3] / g, h[0] / g, h[1] / g, h[2] / g, h[6] / g, h[4] / g, h[5] / g, h[8] / g, h[7] / g, h[10] / g, h[9] / g);
    cudaFree(p->params.prof);
  }
  delete p;
}

// ---- pyramid head (head_tc.cuh) ---------------------------------------------------------------------------
struct HeadPlan {
  HeadParams params;
  int dt, grid;
};

bool head_tc_supported(int dt, int C, int pc) {
  const int ck = 128 / (int)act_size(dt);
  return (pc == 4 || pc == 2) && C % ck == 0 && C / ck <= kHeadMaxChunks;
}

HeadPlan* head_tc_plan_create(int dt, const void* act, const void* w_packed, const float* bias, const float* prev4,
                              float* out4, int B, int H, int W, int C, int pc, int num_sms, char* err, int errlen,
                              const float* aff) {
  if (!head_tc_supported(dt, C, pc)) {
    snprintf(err, errlen, "pyramid head: unsupported C=%d pc=%d", C, pc);
    return nullptr;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled unavailable"); return nullptr; }
  HeadPlan* p = new HeadPlan();
  memset(&p->params, 0, sizeof(p->params));
  p->dt = dt;
  const cuuint64_t es = act_size(dt);
  const cuuint32_t ck = 128 / es;
  const CUtensorMapDataType tdt = dt == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
    cuuint32_t box[4] = {ck, kHeadWin, kHeadWin, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&p->params.tmA, tdt, 4, const_cast<void*>(act), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, errlen, "pyramid head: act tensor map failed: %d", (int)r); delete p; return nullptr; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)kHeadN};
    cuuint64_t strides[1] = {C * es};
    cuuint32_t box[2] = {ck, kHeadN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p->params.tmW, tdt, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, errlen, "pyramid head: weight tensor map failed: %d", (int)r); delete p; return nullptr; }
  }
  HeadParams& P = p->params;
  P.nchunks = C / ck;
  P.B = B; P.H = H; P.W = W;
  P.tiles_w = (W + kHeadTile - 1) / kHeadTile;
  P.tiles_h = (H + kHeadTile - 1) / kHeadTile;
  P.ntiles = P.tiles_w * P.tiles_h * B;
  P.bias = bias; P.out4 = out4; P.prev4 = prev4; P.pc = pc;
  P.aff = aff; P.C = C;
  p->grid = P.ntiles < num_sms ? P.ntiles : num_sms;
  cudaFuncSetAttribute(head_tc_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem);
  cudaFuncSetAttribute(head_tc_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem);
  cudaFuncSetAttribute(head_tc_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem);
  cudaFuncSetAttribute(head_tc_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem);
  return p;
}

void head_tc_plan_destroy(HeadPlan* p) { delete p; }

void head_tc_launch(const HeadPlan* p, cudaStream_t st) {
  const bool fuse = p->params.aff != nullptr;
  const int threads = kHeadThreads + (fuse ? kHeadXfThreads : 0);
  if (p->dt == kBF16) {
    if (fuse) head_tc_kernel<__nv_bfloat16, true><<<p->grid, threads, kHeadSmem, st>>>(p->params);
    else head_tc_kernel<__nv_bfloat16, false><<<p->grid, threads, kHeadSmem, st>>>(p->params);
  } else {
    if (fuse) head_tc_kernel<float, true><<<p->grid, threads, kHeadSmem, st>>>(p->params);
    else head_tc_kernel<float, false><<<p->grid, threads, kHeadSmem, st>>>(p->params);
  }
}

int tc_conv_tiles_per_image(int dt, int N, int H, int W) {
  const int tile_h = (N == 256) ? 16 : 32;  // NSUB = 1 for N = 256, 2 otherwise (see tc_conv_plan_create)
  return ((W + 7) / 8) * ((H + tile_h - 1) / tile_h);
}

void tc_conv_launch(const TcConvPlan* p, cudaStream_t st) {
  void* args[1] = {const_cast<ConvParams*>(&p->params)};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p->grid);
  cfg.blockDim = dim3(p->threads);
  cfg.dynamicSmemBytes = p->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p->mc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  // programmatic dependent launch: the prologue of this kernel may overlap the tail of its predecessor (conv_tc.cuh);
  // small batches only, see pdl_enabled (kernels.h)
  const bool pdl = pdl_enabled(p->params.B);
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && (p->mc == 1 || p->ks > 1)) ? 2 : 1;
  cudaLaunchKernelExC(&cfg, p->kernel, args);
  if (p->tail) tc_conv_launch(p->tail, st);
}

}  // namespace use
