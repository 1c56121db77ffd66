// libbnvmppi.so -- stateless entry points for the rows either side of the MPPI iteration:
// traversability lookup / collision check, batched environment step, risk-map inference, DWA helpers.
#include "bnv_internal.h"

#include <cmath>
#include <cstdint>

#include "aux_kernels.cuh"

namespace {

bool pow2_float(float v) {
  int e;
  return v > 0.0f && std::isfinite(v) && std::frexp(v, &e) == 0.5f;
}

int make_geom(const bnv_grid* g, bnv::GridGeom* out) {
  if (!g) return bnv_fail(BNV_ERR_INVALID, "null grid");
  if (g->grid_size < 1 || g->pitch < g->grid_size) return bnv_fail(BNV_ERR_INVALID, "grid_size %d / pitch %d invalid", g->grid_size, g->pitch);
  if (!(g->resolution > 0.0f)) return bnv_fail(BNV_ERR_INVALID, "resolution must be positive");
  if (!(g->x_min < g->x_max) || !(g->y_min < g->y_max)) return bnv_fail(BNV_ERR_INVALID, "empty map limits");
  out->x_min = g->x_min;
  out->y_min = g->y_min;
  out->x_max = g->x_max;
  out->y_max = g->y_max;
  out->res = g->resolution;
  out->inv_res = 1.0f / g->resolution;
  out->res_pow2 = pow2_float(g->resolution) ? 1 : 0;
  out->fast_grid = (out->res_pow2 && g->x_min == 0.0f && g->y_min == 0.0f) ? 1 : 0;
  return BNV_OK;
}

// Acklam's rational approximation of the standard normal quantile + one Halley step (double precision; |err| < 1e-15).
double norm_ppf(double p) {
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                             1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                             6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                             -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00};
  double x;
  if (p < 0.02425) {
    double q = std::sqrt(-2 * std::log(p));
    x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  } else if (p <= 1 - 0.02425) {
    double q = p - 0.5, r = q * q;
    x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
        (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
  } else {
    double q = std::sqrt(-2 * std::log(1 - p));
    x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  }
  const double e = 0.5 * std::erfc(-x / std::sqrt(2.0)) - p;
  const double u = e * std::sqrt(2 * M_PI) * std::exp(x * x / 2);
  return x - u / (1 + x * u / 2);
}

}  // namespace

int bnv_launch_argmin(const float* costs, int K, const float* actions, const float* rec, int row_len, float* action_out,
                      float* states_out, int* idx_out, cudaStream_t s) {
  bnv::argmin_gather_kernel<<<1, 256, 0, s>>>(costs, K, actions, rec, row_len, action_out, states_out, idx_out);
  BNV_CUDA(cudaGetLastError());
  return BNV_OK;
}

int bnv_launch_dwa_subgoal(const bnv::GridGeom& geom, int G, const float* tau, int pitch, const bnv::Bounds& b,
                           const float* actions, const float* path, int n, const float* state, float lookahead,
                           float* goal_out, cudaStream_t s) {
  bnv::dwa_subgoal_kernel<<<1, 256, 0, s>>>(geom, G, tau, pitch, b, actions, path, n, state, lookahead, goal_out);
  BNV_CUDA(cudaGetLastError());
  return BNV_OK;
}

extern "C" {

int bnv_trav_lookup(const bnv_grid* grid, const float* mean_dev, const float* std_dev, int64_t env_stride,
                    int64_t rows_per_env, const float* pos_dev, int64_t n, int32_t pos_stride, const float* xi_dev,
                    uint64_t seed, uint64_t counter, const uint64_t* counter_dev, float stuck_threshold,
                    float* trav_out_dev, uint8_t* stuck_out_dev, void* stream) {
  bnv::GridGeom geom;
  int rc = make_geom(grid, &geom);
  if (rc != BNV_OK) return rc;
  if (!mean_dev || !pos_dev || (!trav_out_dev && !stuck_out_dev)) return bnv_fail(BNV_ERR_INVALID, "null argument");
  if (n < 0 || pos_stride < 2 || env_stride < 0 || rows_per_env < 0) return bnv_fail(BNV_ERR_INVALID, "bad size argument");
  if (n == 0) return BNV_OK;
  const long long blocks = (n + 255) / 256;
  if (blocks > 0x7FFFFFFFLL) return bnv_fail(BNV_ERR_UNSUPPORTED, "too many positions for one launch");
  bnv::trav_lookup_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      geom, grid->grid_size, mean_dev, std_dev, grid->pitch, env_stride, rows_per_env, pos_dev, n, pos_stride, xi_dev,
      static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), static_cast<unsigned long long>(counter),
      reinterpret_cast<const unsigned long long*>(counter_dev), stuck_threshold, trav_out_dev, stuck_out_dev);
  BNV_CUDA(cudaGetLastError());
  return BNV_OK;
}

int bnv_env_step(const bnv_grid* grid, const float* mean_dev, const float* std_dev, int64_t env_stride,
                 int32_t num_envs, float* states_dev, const float* actions_dev, const float* goals_dev,
                 const float* xi_dev, uint64_t seed, uint64_t counter, const uint64_t* counter_dev, const float u_min[2],
                 const float u_max[2], float delta_t, float goal_threshold, float* reward_out_dev,
                 uint8_t* terminated_out_dev, void* stream) {
  bnv::GridGeom geom;
  int rc = make_geom(grid, &geom);
  if (rc != BNV_OK) return rc;
  if (!mean_dev || !std_dev || !states_dev || !actions_dev || !goals_dev || !u_min || !u_max || !reward_out_dev ||
      !terminated_out_dev)
    return bnv_fail(BNV_ERR_INVALID, "null argument");
  if (num_envs < 1 || env_stride < 0 || !(delta_t > 0.0f)) return bnv_fail(BNV_ERR_INVALID, "bad size argument");
  bnv::Bounds b{u_min[0], u_min[1], u_max[0], u_max[1], delta_t};
  bnv::env_step_kernel<<<(num_envs + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      geom, grid->grid_size, mean_dev, std_dev, grid->pitch, env_stride, num_envs, states_dev, actions_dev, goals_dev,
      xi_dev, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), static_cast<unsigned long long>(counter),
      reinterpret_cast<const unsigned long long*>(counter_dev), b, goal_threshold, reward_out_dev, terminated_out_dev);
  BNV_CUDA(cudaGetLastError());
  return BNV_OK;
}

int bnv_closed_loop_step(const bnv_grid* grid, const float* mean_dev, const float* std_dev, int64_t env_stride,
                         int32_t num_envs, int32_t horizon, float* states_dev, const float* actions_dev,
                         const float* planned_dev, const float* goals_dev, uint64_t seed, uint64_t* counter_dev,
                         const float u_min[2], const float u_max[2], float delta_t, float goal_threshold,
                         float stuck_threshold, float* reward_out_dev, uint8_t* terminated_out_dev,
                         uint8_t* collisions_out_dev, uint8_t* done_dev, int64_t* steps_to_goal_dev, int64_t* step_no_dev,
                         uint64_t* planner_iteration_dev, uint32_t* ticket_dev, void* stream) {
  bnv::GridGeom geom;
  int rc = make_geom(grid, &geom);
  if (rc != BNV_OK) return rc;
  if (!mean_dev || !std_dev || !states_dev || !actions_dev || !planned_dev || !goals_dev || !counter_dev || !u_min ||
      !u_max || !reward_out_dev || !terminated_out_dev || !collisions_out_dev || !done_dev || !steps_to_goal_dev ||
      !step_no_dev || !ticket_dev)
    return bnv_fail(BNV_ERR_INVALID, "null argument");
  if (num_envs < 1 || horizon < 1 || env_stride < 0 || !(delta_t > 0.0f)) return bnv_fail(BNV_ERR_INVALID, "bad size argument");
  bnv::Bounds b{u_min[0], u_min[1], u_max[0], u_max[1], delta_t};
  bnv::closed_loop_step_kernel<<<num_envs, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      geom, grid->grid_size, mean_dev, std_dev, grid->pitch, env_stride, num_envs, horizon, states_dev, actions_dev,
      planned_dev, goals_dev, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32),
      reinterpret_cast<unsigned long long*>(counter_dev), b, goal_threshold, stuck_threshold, reward_out_dev,
      terminated_out_dev, collisions_out_dev, done_dev, reinterpret_cast<long long*>(steps_to_goal_dev),
      reinterpret_cast<long long*>(step_no_dev), reinterpret_cast<unsigned long long*>(planner_iteration_dev), ticket_dev);
  BNV_CUDA(cudaGetLastError());
  return BNV_OK;
}

int bnv_risk_map(int32_t metric, float confidence, int32_t method, const float* mean_dev, const float* std_dev,
                 int64_t n_cells, const float* samples_dev, int32_t num_samples, uint64_t seed, float* risk_out_dev,
                 float* samples_out_dev, void* stream) {
  if (!mean_dev || !risk_out_dev) return bnv_fail(BNV_ERR_INVALID, "null argument");
  if (metric < 0 || metric > 2) return bnv_fail(BNV_ERR_INVALID, "metric %d is not 0 (expected), 1 (var) or 2 (cvar)", metric);
  if (n_cells < 0) return bnv_fail(BNV_ERR_INVALID, "negative n_cells");
  if (n_cells == 0) return BNV_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long blocks = (n_cells + 255) / 256;
  if (metric == bnv::kRiskExpected) {  // traversability_model.py:38-39: the distribution's mean
    bnv::risk_closed_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(mean_dev, nullptr, n_cells, 0.0f, risk_out_dev);
    BNV_CUDA(cudaGetLastError());
    return BNV_OK;
  }
  if (!(confidence >= 0.0f && confidence <= 1.0f)) return bnv_fail(BNV_ERR_INVALID, "confidence must lie in [0, 1]");  // utils.py:28-31
  if (method == BNV_RISK_CLOSED_FORM) {
    if (!std_dev) return bnv_fail(BNV_ERR_INVALID, "std map required");
    if (!(confidence > 0.0f && confidence < 1.0f)) return bnv_fail(BNV_ERR_INVALID, "closed form needs 0 < confidence < 1");
    const double z = norm_ppf(static_cast<double>(confidence));
    double coef = z;
    if (metric == bnv::kRiskCvar) coef = std::exp(-0.5 * z * z) / std::sqrt(2 * M_PI) / (1.0 - static_cast<double>(confidence));
    bnv::risk_closed_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(mean_dev, std_dev, n_cells, static_cast<float>(coef), risk_out_dev);
    BNV_CUDA(cudaGetLastError());
    return BNV_OK;
  }
  if (method != BNV_RISK_MONTE_CARLO) return bnv_fail(BNV_ERR_INVALID, "unknown method %d", method);
  if (num_samples < 1 || num_samples > 32768) return bnv_fail(BNV_ERR_INVALID, "num_samples %d outside [1, 32768]", num_samples);
  if (!samples_dev && !std_dev) return bnv_fail(BNV_ERR_INVALID, "std map required to draw samples");
  const int s_pad = (num_samples + 31) & ~31;  // whole warps of samples per row; the pads hold +inf
  const size_t row_bytes = static_cast<size_t>(s_pad + 1) * sizeof(float);
  int cpc = static_cast<int>(std::min<size_t>(32, (200 * 1024) / row_bytes));
  if (cpc < 1) return bnv_fail(BNV_ERR_UNSUPPORTED, "num_samples too large for shared memory");
  const size_t smem = cpc * row_bytes;
  BNV_CUDA(cudaFuncSetAttribute(bnv::risk_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  const long long ctas = (n_cells + cpc - 1) / cpc;
  if (ctas > 0x7FFFFFFFLL) return bnv_fail(BNV_ERR_UNSUPPORTED, "too many cells for one launch");
  bnv::risk_mc_kernel<<<static_cast<unsigned>(ctas), bnv::kRiskThreads, smem, s>>>(
      mean_dev, std_dev, n_cells, samples_dev, num_samples, s_pad, cpc, metric, confidence, static_cast<uint32_t>(seed),
      static_cast<uint32_t>(seed >> 32), 0u, 0u, risk_out_dev, samples_out_dev);
  BNV_CUDA(cudaGetLastError());
  return BNV_OK;
}

int bnv_dwa_actions(const float* prev_action_dev, const float u_min[2], const float u_max[2], const float a_lim[2],
                    float delta_t, int32_t num_lin_vel, int32_t num_ang_vel, int32_t horizon, float* actions_out_dev,
                    float* controls_out_dev, void* stream) {
  if (!prev_action_dev || !u_min || !u_max || !a_lim || !actions_out_dev || !controls_out_dev)
    return bnv_fail(BNV_ERR_INVALID, "null argument");
  if (num_lin_vel < 1 || num_ang_vel < 1 || horizon < 1) return bnv_fail(BNV_ERR_INVALID, "bad size argument");
  bnv::Bounds b{u_min[0], u_min[1], u_max[0], u_max[1], delta_t};
  const int n = num_lin_vel * num_ang_vel;
  bnv::dwa_actions_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      prev_action_dev, b, a_lim[0], a_lim[1], delta_t, num_lin_vel, num_ang_vel, horizon, actions_out_dev, controls_out_dev);
  BNV_CUDA(cudaGetLastError());
  return BNV_OK;
}

}  // extern "C"
