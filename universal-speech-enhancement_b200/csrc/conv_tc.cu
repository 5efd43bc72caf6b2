// Host side of the tcgen05 convolution: TMA tensor-map construction and launch.
#include <stdio.h>
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
  int dt, N, nsub;
  int grid, threads, smem;
  const void* kernel;
};

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

static bool encode_w(CUtensorMap* m, int dt, const void* base, int taps, int N, int Ctot, char* err, int errlen) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled unavailable"); return false; }
  const cuuint64_t es = act_size(dt);
  const cuuint32_t ck = 128 / es;
  cuuint64_t dims[3] = {(cuuint64_t)Ctot, (cuuint64_t)N, (cuuint64_t)taps};
  cuuint64_t strides[2] = {Ctot * es, (cuuint64_t)N * Ctot * es};
  cuuint32_t box[3] = {ck, (cuuint32_t)N, 1};
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

template <typename T, int N, int NSUB>
static void fill_kernel(TcConvPlan* p) {
  using C = ConvCfg<T, N, NSUB>;
  p->kernel = reinterpret_cast<const void*>(&conv_tc_kernel<T, N, NSUB>);
  p->threads = C::THREADS;
  p->smem = C::SMEM_BYTES;
  p->nsub = NSUB;
  cudaFuncSetAttribute(conv_tc_kernel<T, N, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
}

bool tc_conv_supported(int dt, int N) { return N == 64 || N == 128 || N == 256; }

TcConvPlan* tc_conv_plan_create(int dt, const TcConvDesc& d, int num_sms, char* err, int errlen) {
  if (!tc_conv_supported(dt, d.N)) {
    snprintf(err, errlen, "tcgen05 conv: unsupported C_out=%d", d.N);
    return nullptr;
  }
  TcConvPlan* p = new TcConvPlan();
  memset(&p->params, 0, sizeof(p->params));
  p->dt = dt;
  p->N = d.N;
  if (dt == kBF16) {
    if (d.N == 256) fill_kernel<__nv_bfloat16, 256, 1>(p);
    else if (d.N == 128) fill_kernel<__nv_bfloat16, 128, 2>(p);
    else fill_kernel<__nv_bfloat16, 64, 2>(p);
  } else {
    if (d.N == 256) fill_kernel<float, 256, 1>(p);
    else if (d.N == 128) fill_kernel<float, 128, 2>(p);
    else fill_kernel<float, 64, 2>(p);
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
        !encode_w(&P.seg[i].tmW, dt, s.w, s.taps, d.N, s.Cw_total, err, errlen)) {
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
  p->grid = P.ntiles < num_sms ? P.ntiles : num_sms;
  return p;
}

void tc_conv_plan_destroy(TcConvPlan* p) { delete p; }

void tc_conv_launch(const TcConvPlan* p, cudaStream_t st) {
  void* args[1] = {const_cast<ConvParams*>(&p->params)};
  cudaLaunchKernel(p->kernel, dim3(p->grid), dim3(p->threads), args, p->smem, st);
}

}  // namespace use
