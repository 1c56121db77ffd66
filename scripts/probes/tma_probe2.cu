// Probe 2: TMA 2-D load whose coordinates are computed per thread from a global load (as in rollout_kernel).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../../benchnav_b200/csrc/ptx_sm100.cuh"
using namespace bnv;
struct alignas(64) Params { CUtensorMap map; float* out; const float* state; int bw, bh, G, rho; };

template <int MODE>
__global__ void __launch_bounds__(128, 1) k_dyn(const __grid_constant__ Params P) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float* dst = reinterpret_cast<float*>(smem + 128);
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  float sx = P.state[0], sy = P.state[1];
  int cx = min(max(__float2int_rd(sx * 2.0f), 0), P.G - 1);
  int cy = min(max(__float2int_rd(sy * 2.0f), 0), P.G - 1);
  int ox = max(0, min(cx - P.rho, P.G - P.bw));
  int oy = max(0, min(cy - P.rho, P.G - P.bh));
  if (MODE == 0) {
    if (tid == 0) { mbar_arrive_expect_tx(bar, P.bw * P.bh * 4); tma_load_2d(dst, &P.map, ox, oy, bar); }
  } else if (MODE == 1) {
    ox = __shfl_sync(0xffffffffu, ox, 0); oy = __shfl_sync(0xffffffffu, oy, 0);
    if (tid == 0) { mbar_arrive_expect_tx(bar, P.bw * P.bh * 4); tma_load_2d(dst, &P.map, ox, oy, bar); }
  } else {
    if (tid < 32) {
      unsigned pred;
      asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
      if (pred) { mbar_arrive_expect_tx(bar, P.bw * P.bh * 4); tma_load_2d(dst, &P.map, ox, oy, bar); }
    }
  }
  mbar_wait(bar, 0);
  if (blockIdx.x == 0) for (int i = tid; i < P.bw * P.bh; i += blockDim.x) P.out[i] = dst[i];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  Enc enc = (Enc)fp;
  const int G = 64, pitch = 64;
  std::vector<float> h(G * pitch);
  for (int i = 0; i < G * pitch; ++i) h[i] = (float)i;
  float *d, *out, *st; cudaMalloc(&d, sizeof(float) * G * pitch); cudaMalloc(&out, sizeof(float) * 65536); cudaMalloc(&st, 12);
  float hs[3] = {8.f, 8.f, 0.7f}; cudaMemcpy(st, hs, 12, cudaMemcpyHostToDevice);
  cudaMemcpy(d, h.data(), sizeof(float) * G * pitch, cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 3; ++mode) {
    Params P{}; P.out = out; P.state = st; P.bw = 16; P.bh = 15; P.G = G; P.rho = 7;
    cuuint64_t gdim[2] = {G, G}; cuuint64_t gstr[1] = {pitch * sizeof(float)};
    cuuint32_t box[2] = {16, 15}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&P.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    size_t smem = 60000;
    cudaFuncSetAttribute(k_dyn<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    cudaFuncSetAttribute(k_dyn<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    cudaFuncSetAttribute(k_dyn<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    cudaMemset(out, 0, sizeof(float) * 65536);
    if (mode == 0) k_dyn<0><<<32, 128, smem>>>(P);
    if (mode == 1) k_dyn<1><<<32, 128, smem>>>(P);
    if (mode == 2) k_dyn<2><<<32, 128, smem>>>(P);
    cudaError_t e = cudaDeviceSynchronize();
    float got[2] = {-1, -1};
    if (e == cudaSuccess) cudaMemcpy(got, out, 8, cudaMemcpyDeviceToHost);
    printf("mode %d: encode=%d run=%s first=%g (want %g)\n", mode, (int)r, cudaGetErrorString(e), got[0], (float)(9 * pitch + 9));
    if (e != cudaSuccess) { printf("context dead; stopping\n"); return 1; }
  }
  return 0;
}
