// Instantiations of the WIDE (throughput) variant of the rollout kernel for the single deterministic solver:
// rollout_kernel<kPatch, kPow2, kRecord, true, kPhilox, false, false, true>.  The wide variant requires the branch-free
// angle path; solvers with dt * max|omega| >= 3 rad per step stay on the latency variant at any size.
#define BNV_ROLLOUT_ONLY
#include "bnv_internal.h"
#include "mppi_kernels.cuh"

namespace {
template <bool R>
BnvRolloutFn pick3(bool patch, bool pow2, bool philox) {
  using namespace bnv;
  if (patch) {
    if (pow2) return philox ? rollout_kernel<true, true, R, true, true, false, false, true> : rollout_kernel<true, true, R, true, false, false, false, true>;
    return philox ? rollout_kernel<true, false, R, true, true, false, false, true> : rollout_kernel<true, false, R, true, false, false, false, true>;
  }
  if (pow2) return philox ? rollout_kernel<false, true, R, true, true, false, false, true> : rollout_kernel<false, true, R, true, false, false, false, true>;
  return philox ? rollout_kernel<false, false, R, true, true, false, false, true> : rollout_kernel<false, false, R, true, false, false, false, true>;
}
}  // namespace

BnvRolloutFn bnv_pick_rollout_wide(bool patch, bool pow2, bool record, bool philox) {
  return record ? pick3<true>(patch, pow2, philox) : pick3<false>(patch, pow2, philox);
}
