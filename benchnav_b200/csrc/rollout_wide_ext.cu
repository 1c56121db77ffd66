// Instantiations of the WIDE (throughput) variant of the rollout kernel for the stochastic-slip and
// batched-environment modes (BASELINE configs 4 and 3): rollout_kernel<kPatch, kPow2, true, true, kPhilox, S, B, true>.
#define BNV_ROLLOUT_ONLY
#include "bnv_internal.h"
#include "mppi_kernels.cuh"

namespace {
template <bool S, bool B>
BnvRolloutFn pick3(bool patch, bool pow2, bool philox) {
  using namespace bnv;
  if (patch) {
    if (pow2) return philox ? rollout_kernel<true, true, true, true, true, S, B, true> : rollout_kernel<true, true, true, true, false, S, B, true>;
    return philox ? rollout_kernel<true, false, true, true, true, S, B, true> : rollout_kernel<true, false, true, true, false, S, B, true>;
  }
  if (pow2) return philox ? rollout_kernel<false, true, true, true, true, S, B, true> : rollout_kernel<false, true, true, true, false, S, B, true>;
  return philox ? rollout_kernel<false, false, true, true, true, S, B, true> : rollout_kernel<false, false, true, true, false, S, B, true>;
}
}  // namespace

BnvRolloutFn bnv_pick_rollout_wide_ext(bool patch, bool pow2, bool philox, bool stoch, bool batch) {
  if (stoch && batch) return pick3<true, true>(patch, pow2, philox);
  if (stoch) return pick3<true, false>(patch, pow2, philox);
  if (batch) return pick3<false, true>(patch, pow2, philox);
  return nullptr;
}
