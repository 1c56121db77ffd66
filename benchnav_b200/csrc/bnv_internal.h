// Shared by the translation units of libbnvmppi.so: error reporting and the rollout-kernel dispatch.
#pragma once
#include <cuda_runtime.h>

#include "../../include/bnv_mppi.h"

namespace bnv {
struct EngineParams;
struct GridGeom;
struct Bounds;
}

// Records a thread-local message for bnv_last_error() and returns `code` (defined in bnv_mppi.cu).
int bnv_fail(int code, const char* fmt, ...);

#define BNV_CUDA(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return bnv_fail(BNV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

using BnvRolloutFn = void (*)(bnv::EngineParams);

// rollout_kernel<kPatch, kPow2, true, true, kPhilox, kStoch, kBatch> for the stochastic / batched modes
// (instantiated in rollout_ext.cu so that the two halves of the template space compile in parallel).
BnvRolloutFn bnv_pick_rollout_ext(bool patch, bool pow2, bool philox, bool stoch, bool batch);

// the wide (throughput) variant, rollout_kernel<..., kWide = true> (rollout_wide.cu / rollout_wide_ext.cu)
BnvRolloutFn bnv_pick_rollout_wide(bool patch, bool pow2, bool record, bool philox);
BnvRolloutFn bnv_pick_rollout_wide_ext(bool patch, bool pow2, bool philox, bool stoch, bool batch);

// argmin_gather_kernel (aux_kernels.cuh) launcher, defined in bnv_aux.cu.
int bnv_launch_argmin(const float* costs, int K, const float* actions, const float* rec, int row_len, float* action_out,
                      float* states_out, int* idx_out, cudaStream_t s);

// dwa_subgoal_kernel (aux_kernels.cuh) launcher, defined in bnv_aux.cu.
int bnv_launch_dwa_subgoal(const bnv::GridGeom& geom, int G, const float* tau, int pitch, const bnv::Bounds& b,
                           const float* actions, const float* path, int n, const float* state, float lookahead,
                           float* goal_out, cudaStream_t s);
