// Host side of the tcgen05 convolution: TMA tensor-map construction and launch.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "conv_tc.cuh"
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
  int dt, N, nsub, mc;
  int grid, threads, smem;
  const void* kernel;
};

// CTA-pair weight multicast (cluster of 2) is OFF by default: measured on B200 it does not pay (bf16, batch 8:
// 25.3 ms of convolutions per evaluation with it, 24.0 ms without) because the C_out = 128 layers are bound by the
// shared-memory read bandwidth of single-CTA M128xN128 MMAs, not by L2 -> SM weight traffic.  USE_B200_CONV_MC=2
// enables it for A/B measurements.
static int multicast_width() {
  static int mc = [] {
    const char* v = getenv("USE_B200_CONV_MC");
    return (v && v[0] == '2') ? 2 : 1;
  }();
  return mc;
}

static bool encode_act(CUtensorMap* m, int dt, const void* base, int B, int H, int W, int Ct, int rows, char* err,
                       int errlen) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled unavailable"); return false; }
  const cuuint64_t es = act_size(dt);
  const cuuint32_t ck = 128 / es;
  cuuint64_t dims[4] = {(cuuint64_t)Ct, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {Ct * es, (cuuint64_t)W * Ct * es, (cuuint64_t)H * W * Ct * es};
  cuuint32_t box[4] = {ck, 8, (cuuint32_t)rows, 1};
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

// swap-AB variant of the C_out = 128 kernel (see conv_tc.cuh); USE_B200_CONV_SWAP=0 selects the pixel-major one.
static bool swap_ab_enabled() {
  static bool on = [] {
    const char* v = getenv("USE_B200_CONV_SWAP");
    return !(v && v[0] == '0');
  }();
  return on;
}

template <typename T, int N, int NSUB>
static void fill_kernel(TcConvPlan* p) {
  using C = ConvCfg<T, N, NSUB>;
  if (p->mc == 2) {
    p->kernel = reinterpret_cast<const void*>(&conv_tc_kernel<T, N, NSUB, 2, false>);
    cudaFuncSetAttribute(conv_tc_kernel<T, N, NSUB, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  } else if (N == 128 && NSUB == 2 && swap_ab_enabled()) {
    if constexpr (N == 128 && NSUB == 2) {
      p->kernel = reinterpret_cast<const void*>(&conv_tc_kernel<T, N, NSUB, 1, true>);
      cudaFuncSetAttribute(conv_tc_kernel<T, N, NSUB, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    }
  } else {
    p->kernel = reinterpret_cast<const void*>(&conv_tc_kernel<T, N, NSUB, 1, false>);
    cudaFuncSetAttribute(conv_tc_kernel<T, N, NSUB, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  }
  p->threads = C::THREADS;
  p->smem = C::SMEM_BYTES;
  p->nsub = NSUB;
}

bool tc_conv_supported(int dt, int N) { return N == 32 || N == 64 || N == 128 || N == 256; }

TcConvPlan* tc_conv_plan_create(int dt, const TcConvDesc& d, int num_sms, char* err, int errlen) {
  if (!tc_conv_supported(dt, d.N) || ((d.N == 32) != (d.out4 != nullptr))) {
    snprintf(err, errlen, "tcgen05 conv: unsupported C_out=%d (out4 %s)", d.N, d.out4 ? "set" : "unset");
    return nullptr;
  }
  TcConvPlan* p = new TcConvPlan();
  memset(&p->params, 0, sizeof(p->params));
  p->dt = dt;
  p->N = d.N;
  p->mc = (d.N >= 64) ? multicast_width() : 1;  // the 32-wide pyramid head has 4 KB weight tiles: not worth pairing
  if (dt == kBF16) {
    if (d.N == 256) fill_kernel<__nv_bfloat16, 256, 1>(p);
    else if (d.N == 128) fill_kernel<__nv_bfloat16, 128, 2>(p);
    else if (d.N == 64) fill_kernel<__nv_bfloat16, 64, 2>(p);
    else fill_kernel<__nv_bfloat16, 32, 2>(p);
  } else {
    if (d.N == 256) fill_kernel<float, 256, 1>(p);
    else if (d.N == 128) fill_kernel<float, 128, 2>(p);
    else if (d.N == 64) fill_kernel<float, 64, 2>(p);
    else fill_kernel<float, 32, 2>(p);
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
    const int rows = s.taps == 9 ? tile_h + 2 : tile_h;
    if (!encode_act(&P.seg[i].tmA, dt, s.act, d.B, d.H, d.W, s.C_tensor, rows, err, errlen) ||
        !encode_w(&P.seg[i].tmW, dt, s.w, s.taps, d.N, s.Cw_total, d.N, err, errlen) ||
        !encode_w(&P.seg[i].tmWh, dt, s.w, s.taps, d.N, s.Cw_total, d.N / 2, err, errlen)) {
      delete p;
      return nullptr;
    }
    P.seg[i].nchunks = s.C / ck;
    P.seg[i].taps = s.taps;
    P.seg[i].wc0 = s.wc0;
    P.seg[i].ac0 = s.c0;
  }
  P.B = d.B; P.H = d.H; P.W = d.W;
  P.tiles_w = (d.W + 7) / 8;
  P.tiles_h = (d.H + tile_h - 1) / tile_h;
  P.ntiles = P.tiles_w * P.tiles_h * d.B;
  P.out = d.out; P.bias = d.bias; P.bias_bstride = d.bias_bstride; P.res = d.res; P.scale = d.scale;
  P.stats_acc = d.stats_acc;
  P.out4 = d.out4; P.prev4 = d.prev4; P.out_pc = d.out_pc ? d.out_pc : 4;
  const int units = (P.ntiles + p->mc - 1) / p->mc;           // tile groups
  const int max_groups = num_sms / p->mc;
  p->grid = (units < max_groups ? units : max_groups) * p->mc;  // a multiple of the cluster width
  return p;
}

void tc_conv_plan_destroy(TcConvPlan* p) { delete p; }

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
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p->mc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelExC(&cfg, p->kernel, args);
}

}  // namespace use
