// Microarchitecture probe (not product code): does a tcgen05 SWIZZLE_128B K-major shared-memory descriptor whose start
// address is offset by whole 128-byte rows (not 1024-byte aligned), and/or whose 8-row groups are SBO = 1280 B apart,
// read the rows it points at when every row was written with the ABSOLUTE-address swizzle (16-byte chunk j of the row
// at byte address a lives at a + ((j ^ ((a >> 7) & 7)) << 4))?  If yes, a 3x3 convolution can stage ONE
// (rows+2) x 10-pixel activation window per channel chunk and address all nine taps inside it, instead of three
// horizontally shifted 8-pixel-wide copies.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../universal-speech-enhancement_b200/csrc \
//        swz_probe.cu -o swz_probe -cudart static && ./swz_probe
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "common.cuh"

using namespace use;

__device__ __forceinline__ uint64_t desc_custom(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

constexpr int kRows = 400;  // pixel rows in the probe buffer
constexpr int kN = 64;

__host__ __device__ inline float fval(int p, int c) { return static_cast<float>((p * 7 + c * 3) % 11 - 5); }
__host__ __device__ inline float wval(int n, int c) { return static_cast<float>((n * 5 + c) % 7 - 3); }

// mode 0: bf16 (K chunk = 64 channels); mode 1: tf32 (32 channels)
template <bool kBf16>
__global__ void __launch_bounds__(128, 1) probe_kernel(int off_rows, int sbo, int bo_mode, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem;                      // kRows x 128 B
  uint8_t* sW = smem + ((kRows * 128 + 1023) & ~1023);  // kN x 128 B, 1024-aligned, standard layout
  uint64_t* bar = reinterpret_cast<uint64_t*>(sW + kN * 128);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
  constexpr int CK = kBf16 ? 64 : 32;
  constexpr int EPV = kBf16 ? 8 : 4;  // elements per 16-byte chunk
  for (int i = threadIdx.x; i < kRows * 8; i += blockDim.x) {
    const int p = i >> 3, j = i & 7;
    const uint32_t a = smem_u32(sP) + p * 128;
    uint8_t* dst = sP + p * 128 + ((j ^ ((a >> 7) & 7)) << 4);
    if constexpr (kBf16) {
      __nv_bfloat16 v[8];
      for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16(fval(p, j * EPV + e));
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(v);
    } else {
      float v[4];
      for (int e = 0; e < 4; ++e) v[e] = fval(p, j * EPV + e);
      *reinterpret_cast<float4*>(dst) = *reinterpret_cast<float4*>(v);
    }
  }
  for (int i = threadIdx.x; i < kN * 8; i += blockDim.x) {
    const int n = i >> 3, j = i & 7;
    uint8_t* dst = sW + n * 128 + ((j ^ (n & 7)) << 4);
    if constexpr (kBf16) {
      __nv_bfloat16 v[8];
      for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16(wval(n, j * EPV + e));
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(v);
    } else {
      float v[4];
      for (int e = 0; e < 4; ++e) v[e] = wval(n, j * EPV + e);
      *reinterpret_cast<float4*>(dst) = *reinterpret_cast<float4*>(v);
    }
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(tslot, 64); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tslot;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(sP) + off_rows * 128;
    const uint32_t bo = bo_mode ? ((a0 >> 7) & 7) : 0;
    constexpr uint32_t idesc = umma_idesc(kBf16 ? 1 : 2, 128, kN);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t ad = desc_custom(a0 + k * 32, sbo, bo);
      const uint64_t bd = desc_custom(smem_u32(sW) + k * 32, 1024, 0);
      umma_ss<kBf16>(tb, ad, bd, idesc, k ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < kN; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tb + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * kN + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 64);
  (void)CK;
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * kN * 4);
  std::vector<float> h(128 * kN);
  const size_t sm = 1024 + ((kRows * 128 + 1023) & ~1023) + kN * 128 + 64;
  cudaFuncSetAttribute(probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaFuncSetAttribute(probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  const int offs[] = {0, 1, 3, 8, 11, 21};
  const int sbos[] = {1024, 1280};
  for (int mode = 0; mode < 2; ++mode)
    for (int sbo : sbos)
      for (int off : offs)
        for (int bo = 0; bo < 2; ++bo) {
          cudaMemset(d_out, 0, 128 * kN * 4);
          if (mode == 0) probe_kernel<true><<<1, 128, sm>>>(off, sbo, bo, d_out);
          else probe_kernel<false><<<1, 128, sm>>>(off, sbo, bo, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h.data(), d_out, 128 * kN * 4, cudaMemcpyDeviceToHost);
          const int CK = mode == 0 ? 64 : 32;
          int bad = 0;
          double maxerr = 0;
          for (int m = 0; m < 128; ++m) {
            const int p = off + (m / 8) * (sbo / 128) + (m % 8);
            for (int n = 0; n < kN; ++n) {
              double ref = 0;
              for (int c = 0; c < CK; ++c) ref += (double)fval(p, c) * wval(n, c);
              const double err = fabs(ref - h[m * kN + n]);
              if (err > 1e-3) ++bad;
              if (err > maxerr) maxerr = err;
            }
          }
          printf("PROBE mode=%s sbo=%d off_rows=%d base_offset=%s : %s (bad=%d maxerr=%.1f)\n", mode == 0 ? "bf16" : "tf32",
                 sbo, off, bo ? "formula" : "0", bad == 0 ? "OK" : "WRONG", bad, maxerr);
        }
  return 0;
}
