// Standalone probe: which way of handing a 2-D tensor map to UTMALDG works on this driver/GPU?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../benchnav_b200/csrc/ptx_sm100.cuh"
using namespace bnv;

struct alignas(64) Params { CUtensorMap map; float* out; int bw, bh, ox, oy; };

__global__ void k_param(const __grid_constant__ Params P) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float* dst = reinterpret_cast<float*>(smem + 128);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_arrive_expect_tx(bar, P.bw * P.bh * 4); tma_load_2d(dst, &P.map, P.ox, P.oy, bar); }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < P.bw * P.bh; i += blockDim.x) P.out[i] = dst[i];
}
__global__ void k_global(const CUtensorMap* map, float* out, int bw, int bh, int ox, int oy) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float* dst = reinterpret_cast<float*>(smem + 128);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_arrive_expect_tx(bar, bw * bh * 4); tma_load_2d(dst, map, ox, oy, bar); }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = dst[i];
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
  float *d, *out; cudaMalloc(&d, sizeof(float) * G * pitch); cudaMalloc(&out, sizeof(float) * 65536);
  cudaMemcpy(d, h.data(), sizeof(float) * G * pitch, cudaMemcpyHostToDevice);
  int boxes[4][2] = {{28, 25}, {32, 32}, {64, 16}, {4, 1}};
  for (auto& b : boxes) {
    for (int variant = 0; variant < 2; ++variant) {
      Params P{}; P.out = out; P.bw = b[0]; P.bh = b[1]; P.ox = 4; P.oy = 3;
      cuuint64_t gdim[2] = {G, G}; cuuint64_t gstr[1] = {pitch * sizeof(float)};
      cuuint32_t box[2] = {(cuuint32_t)b[0], (cuuint32_t)b[1]}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&P.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      size_t smem = 128 + b[0] * b[1] * 4;
      cudaMemset(out, 0, sizeof(float) * 65536);
      if (variant == 0) {
        k_param<<<1, 128, smem>>>(P);
      } else {
        CUtensorMap* dm; cudaMalloc(&dm, sizeof(CUtensorMap)); cudaMemcpy(dm, &P.map, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        k_global<<<1, 128, smem>>>(dm, out, b[0], b[1], 4, 3);
      }
      cudaError_t e = cudaDeviceSynchronize();
      float got[2] = {-1, -1};
      if (e == cudaSuccess) cudaMemcpy(got, out, 8, cudaMemcpyDeviceToHost);
      printf("box %dx%d variant %s: encode=%d run=%s first=%g (want %g) second=%g\n", b[0], b[1], variant ? "global" : "param",
             (int)r, cudaGetErrorString(e), got[0], (float)(3 * pitch + 4), got[1]);
      if (e != cudaSuccess) { printf("context dead; stopping\n"); return 1; }
    }
  }
  return 0;
}
