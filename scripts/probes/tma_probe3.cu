// Probe 3: one TMA 2-D load per process; args: bw bh ox oy nblocks smem_bytes G
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../benchnav_b200/csrc/ptx_sm100.cuh"
using namespace bnv;
struct alignas(64) Params { CUtensorMap map; float* out; int bw, bh, ox, oy; };
__global__ void k(const __grid_constant__ Params P) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float* dst = reinterpret_cast<float*>(smem + 128);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x < 32) {
    if (elect_one()) { mbar_arrive_expect_tx(bar, P.bw * P.bh * 4); tma_load_2d(dst, &P.map, P.ox, P.oy, bar); }
  }
  mbar_wait(bar, 0);
  if (blockIdx.x == 0) for (int i = threadIdx.x; i < P.bw * P.bh; i += blockDim.x) P.out[i] = dst[i];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  int bw = atoi(argv[1]), bh = atoi(argv[2]), ox = atoi(argv[3]), oy = atoi(argv[4]), nb = atoi(argv[5]);
  size_t smem = atol(argv[6]); int G = atoi(argv[7]);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  Enc enc = (Enc)fp;
  int pitch = (G + 3) & ~3;
  std::vector<float> h(G * pitch);
  for (int i = 0; i < G * pitch; ++i) h[i] = (float)i;
  float *d, *out; cudaMalloc(&d, sizeof(float) * G * pitch); cudaMalloc(&out, sizeof(float) * 65536);
  cudaMemcpy(d, h.data(), sizeof(float) * G * pitch, cudaMemcpyHostToDevice);
  Params P{}; P.out = out; P.bw = bw; P.bh = bh; P.ox = ox; P.oy = oy;
  cuuint64_t gdim[2] = {(cuuint64_t)G, (cuuint64_t)G}; cuuint64_t gstr[1] = {pitch * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&P.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
  cudaMemset(out, 0, sizeof(float) * 65536);
  k<<<nb, 128, smem>>>(P);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> got(bw * bh, -1.f);
  int bad = -1;
  if (e == cudaSuccess) {
    cudaMemcpy(got.data(), out, sizeof(float) * bw * bh, cudaMemcpyDeviceToHost);
    bad = 0;
    for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
      float want = (ox + x >= 0 && ox + x < G && oy + y >= 0 && oy + y < G) ? (float)((oy + y) * pitch + ox + x) : 0.f;
      if (got[y * bw + x] != want) ++bad;
    }
  }
  printf("box %dx%d at (%d,%d) blocks %d smem %zu G %d: encode=%d run=%s mismatches=%d\n", bw, bh, ox, oy, nb, smem, G, (int)r,
         cudaGetErrorString(e), bad);
  return 0;
}
