// sm_100a kernels for the rows either side of the MPPI iteration:
//
//   trav_lookup_kernel     TraversabilityModel.get_traversability for arbitrary positions, inference or observation
//                          mode, + the `trav <= threshold` test of PlanetaryEnv.collision_check
//                          (traversability_model.py:53-72, planetary_env.py:221-232)
//   env_step_kernel        PlanetaryEnv.step for E independent environments (planetary_env.py:189-219)
//   risk_closed_kernel     risk map, closed form for a Normal slip distribution (expected value / VaR / CVaR)
//   risk_mc_kernel         risk map by Monte-Carlo exactly as TraversabilityModel._infer_risk_map
//                          (traversability_model.py:28-51): S draws per cell, torch.quantile (linear) by radix
//                          selection of the two order statistics, tail nanmean
//   dwa_actions_kernel     DWA._generate_actions (dwa.py:151-184): dynamic window, linspace, cartesian product
//   dwa_subgoal_kernel     DWA._select_sub_goal (dwa.py:260-285)
//   argmin_gather_kernel   DWA.forward's argmin + gathers (dwa.py:141-144)
//
// All of these are small, HBM/latency-bound helpers; none is on the K x T critical path.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "mppi_math.cuh"

namespace bnv {

// Stream ids (Philox counter word 1) of the helpers' draws; the MPPI rollout uses word 1 = step pair (< 2^31) for
// the control noise and 0x80000000 | pair for its lookups, keyed by the solver seed -- helpers take their own seed.
constexpr uint32_t kStreamLookup = 0x40000000u;
constexpr uint32_t kStreamEnvStep = 0x40000001u;
constexpr uint32_t kStreamRisk = 0x20000000u;  // | quad index

__device__ __forceinline__ float aux_normal(uint32_t index, uint32_t stream, uint32_t ctr_lo, uint32_t ctr_hi, uint2 key) {
  const uint4 r = philox4x32_10(make_uint4(index, stream, ctr_lo, ctr_hi), key);
  return box_muller(r.x, r.y).x;
}

// Cell of a position, clamped to the map (grid_map.py:195-209).
__device__ __forceinline__ int cell_of(const GridGeom& g, int G, int pitch, float x, float y) {
  const int ix = min(max(cell_coord_rt(x, g.x_min, g), 0), G - 1);
  const int iy = min(max(cell_coord_rt(y, g.y_min, g), 0), G - 1);
  return iy * pitch + ix;
}

// ------------------------------------------------------------------------------------------------ lookup / collision
// pos: n rows of `pos_stride` floats, (x, y) first.  stdv == nullptr: inference mode on a risk map `mean`
// (traversability_model.py:70-72).  Otherwise observation mode: 1 - clamp(Normal(mean, std).sample(), 0, 1)
// (traversability_model.py:65-69) with the standard normal injected (xi[n]) or drawn from Philox.
// `env_rows` > 0: positions [E][env_rows] index E stacked maps (env_stride elements apart).
__global__ void __launch_bounds__(256) trav_lookup_kernel(GridGeom geom, int G, const float* __restrict__ mean,
                                                          const float* __restrict__ stdv, int pitch,
                                                          long long env_stride, long long env_rows,
                                                          const float* __restrict__ pos, long long n, int pos_stride,
                                                          const float* __restrict__ xi, uint32_t seed_lo,
                                                          uint32_t seed_hi, unsigned long long ctr,
                                                          const unsigned long long* __restrict__ ctr_dev, float thr,
                                                          float* __restrict__ trav_out,
                                                          unsigned char* __restrict__ stuck_out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (ctr_dev != nullptr) ctr += *ctr_dev;  // device-resident draw counter (graph-capturable callers)
  const uint32_t ctr_lo = static_cast<uint32_t>(ctr), ctr_hi = static_cast<uint32_t>(ctr >> 32);
  const float x = pos[i * pos_stride], y = pos[i * pos_stride + 1];
  const size_t cell = static_cast<size_t>(cell_of(geom, G, pitch, x, y)) +
                      (env_rows > 0 ? static_cast<size_t>(i / env_rows) * env_stride : 0);
  float tau;
  if (stdv == nullptr) {
    const float r = mean[cell];
    const float c = (r != r) ? r : fminf(fmaxf(r, 0.0f), 1.0f);
    tau = __fsub_rn(1.0f, c);
  } else {
    const float z = xi ? xi[i]
                       : aux_normal(static_cast<uint32_t>(i), kStreamLookup + static_cast<uint32_t>(i >> 32) * 2u, ctr_lo,
                                    ctr_hi, make_uint2(seed_lo, seed_hi));
    tau = slip_to_trav(make_float2(mean[cell], stdv[cell]), z);
  }
  if (trav_out) trav_out[i] = tau;
  if (stuck_out) stuck_out[i] = (tau <= thr) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ environment step
// One thread = one environment: observation-mode transit of the robot state (robot_model.py:75-98 with the
// environment's delta_t), reward = the traversability drawn for the step, terminated = ||p - goal|| < goal_threshold
// (planetary_env.py:203-217).  states [E][3] are updated in place; maps are E stacked [G][pitch] arrays
// (env_stride = 0: one shared map).
__global__ void __launch_bounds__(128) env_step_kernel(GridGeom geom, int G, const float* __restrict__ mean,
                                                       const float* __restrict__ stdv, int pitch, long long env_stride,
                                                       int E, float* __restrict__ states,
                                                       const float* __restrict__ actions,
                                                       const float* __restrict__ goals, const float* __restrict__ xi,
                                                       uint32_t seed_lo, uint32_t seed_hi, unsigned long long ctr,
                                                       const unsigned long long* __restrict__ ctr_dev, Bounds b,
                                                       float goal_threshold, float* __restrict__ reward,
                                                       unsigned char* __restrict__ terminated) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  if (ctr_dev != nullptr) ctr += *ctr_dev;  // device-resident draw counter (graph-capturable callers)
  const uint32_t ctr_lo = static_cast<uint32_t>(ctr), ctr_hi = static_cast<uint32_t>(ctr >> 32);
  float x = states[3 * e], y = states[3 * e + 1], th = states[3 * e + 2];
  const size_t cell = static_cast<size_t>(cell_of(geom, G, pitch, x, y)) + static_cast<size_t>(e) * env_stride;
  const float z = xi ? xi[e] : aux_normal(static_cast<uint32_t>(e), kStreamEnvStep, ctr_lo, ctr_hi, make_uint2(seed_lo, seed_hi));
  const float tau = slip_to_trav(make_float2(mean[cell], stdv[cell]), z);
  const float v0 = clampf(actions[2 * e], b.u_min0, b.u_max0);
  const float v1 = clampf(actions[2 * e + 1], b.u_min1, b.u_max1);
  StepConsts c{};
  c.x_min = geom.x_min; c.y_min = geom.y_min; c.x_max = geom.x_max; c.y_max = geom.y_max; c.dt = b.dt;
  float xr, yr, thr;
  unicycle_step<false>(c, tau, v0, v1, x, y, th, xr, yr, thr);
  states[3 * e] = x;
  states[3 * e + 1] = y;
  states[3 * e + 2] = th;
  reward[e] = tau;
  const float dx = __fsub_rn(x, goals[2 * e]), dy = __fsub_rn(y, goals[2 * e + 1]);
  const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  terminated[e] = (d < goal_threshold) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ closed-loop step
// Everything Tutorial 3.3's loop does between two planner calls (test/test_mppi.py:180-185), for E environments in ONE
// launch: apply the first planned action to the environment (PlanetaryEnv.step, planetary_env.py:189-219; robots that
// have arrived stop), test the planned trajectory for collisions (PlanetaryEnv.collision_check, :221-232), and keep
// the loop's books (step counter, first step at which each robot reached its goal, draw counters) -- so that a
// control step is two kernels (rollout + this one) instead of the ~15 small launches the same sequence costs through
// separate calls.  blockIdx.x = environment; thread 0 advances the robot, threads 0..T test the planned states.
// Draws: exactly those of env_step_kernel (counter c) and trav_lookup_kernel (counter 2^40 + c + 1) called one after
// the other, so the fused loop reproduces the unfused one bit for bit; the last block to finish advances *ctr_dev by 2,
// *step_no by 1 and, when given, the planner's iteration counter by 1 (the rollout kernel then needs no bump kernel).
__global__ void __launch_bounds__(128) closed_loop_step_kernel(
    GridGeom geom, int G, const float* __restrict__ mean, const float* __restrict__ stdv, int pitch, long long env_stride,
    int E, int T, float* __restrict__ states, const float* __restrict__ actions /*[E][T][2]*/,
    const float* __restrict__ planned /*[E][T+1][3]*/, const float* __restrict__ goals, uint32_t seed_lo,
    uint32_t seed_hi, unsigned long long* __restrict__ ctr_dev, Bounds b, float goal_threshold, float stuck_threshold,
    float* __restrict__ reward, unsigned char* __restrict__ terminated, unsigned char* __restrict__ collisions,
    unsigned char* __restrict__ done, long long* __restrict__ steps_to_goal, long long* __restrict__ step_no,
    unsigned long long* __restrict__ planner_iter, unsigned int* __restrict__ ticket) {
  const int e = blockIdx.x, tid = threadIdx.x;
  const unsigned long long ctr = *ctr_dev;
  const uint2 key = make_uint2(seed_lo, seed_hi);
  // planned trajectory against the environment's true slip model (counter 2^40 + c + 1, index e (T+1) + t)
  for (int t = tid; t <= T; t += blockDim.x) {
    const long long i = static_cast<long long>(e) * (T + 1) + t;
    const float* p = planned + i * 3;
    const size_t cell = static_cast<size_t>(cell_of(geom, G, pitch, p[0], p[1])) + static_cast<size_t>(e) * env_stride;
    const unsigned long long cc = (1ull << 40) + ctr + 1ull;
    const float z = aux_normal(static_cast<uint32_t>(i), kStreamLookup + static_cast<uint32_t>(i >> 32) * 2u,
                               static_cast<uint32_t>(cc), static_cast<uint32_t>(cc >> 32), key);
    collisions[i] = (slip_to_trav(make_float2(mean[cell], stdv[cell]), z) <= stuck_threshold) ? 1 : 0;
  }
  if (tid == 0) {
    const bool was_done = done[e] != 0;
    float x = states[3 * e], y = states[3 * e + 1], th = states[3 * e + 2];
    const size_t cell = static_cast<size_t>(cell_of(geom, G, pitch, x, y)) + static_cast<size_t>(e) * env_stride;
    const float z = aux_normal(static_cast<uint32_t>(e), kStreamEnvStep, static_cast<uint32_t>(ctr),
                               static_cast<uint32_t>(ctr >> 32), key);
    const float tau = slip_to_trav(make_float2(mean[cell], stdv[cell]), z);
    const float a0 = was_done ? 0.0f : actions[static_cast<size_t>(e) * T * 2];
    const float a1 = was_done ? 0.0f : actions[static_cast<size_t>(e) * T * 2 + 1];
    const float v0 = clampf(a0, b.u_min0, b.u_max0), v1 = clampf(a1, b.u_min1, b.u_max1);
    StepConsts c{};
    c.x_min = geom.x_min; c.y_min = geom.y_min; c.x_max = geom.x_max; c.y_max = geom.y_max; c.dt = b.dt;
    float xr, yr, thr;
    unicycle_step<false>(c, tau, v0, v1, x, y, th, xr, yr, thr);
    states[3 * e] = x;
    states[3 * e + 1] = y;
    states[3 * e + 2] = th;
    reward[e] = tau;
    const float dx = __fsub_rn(x, goals[2 * e]), dy = __fsub_rn(y, goals[2 * e + 1]);
    const bool term = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))) < goal_threshold;
    terminated[e] = term ? 1 : 0;
    if (term && !was_done) steps_to_goal[e] = *step_no + 1;
    if (term) done[e] = 1;
  }
  // the counters advance once every block has read them: last block out (release/acquire through the ticket)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0u;
      *ctr_dev = ctr + 2ull;
      *step_no += 1;
      if (planner_iter != nullptr) *planner_iter += 1ull;
    }
  }
}

// ------------------------------------------------------------------------------------------------ risk map
// Closed form for a Normal(mean, std) slip distribution: risk = mean + coef * std with
//   expected value: coef = 0;  VaR_q: coef = Phi^-1(q);  CVaR_q: coef = phi(Phi^-1(q)) / (1 - q)
// (coef computed on the host in double precision).  The reference estimates the same quantities from 1000 draws per
// cell (traversability_model.py:40-51); this is the limit of that estimator, without its Monte-Carlo error.
__global__ void __launch_bounds__(256) risk_closed_kernel(const float* __restrict__ mean, const float* __restrict__ stdv,
                                                          long long n, float coef, float* __restrict__ risk) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) risk[i] = (coef == 0.0f) ? mean[i] : fmaf(coef, stdv[i], mean[i]);
}

constexpr int kRiskThreads = 256;
enum RiskMetric { kRiskExpected = 0, kRiskVar = 1, kRiskCvar = 2 };

// Monte-Carlo risk map, the reference's estimator: per cell S samples (injected [S][n] exactly as
// `distributions.sample((S,))`, or drawn here: sample = z * std + mean), VaR = torch.quantile(samples, q, dim=0)
// (linear interpolation between order statistics floor/ceil(q (S-1)), ATen lerp), CVaR = mean of the samples
// strictly above VaR (nanmean of the masked tail; NaN when the tail is empty).
// One CTA = `cpc` consecutive cells, samples staged in shared memory as rows of S_pad (+1 pad) floats, +inf padded;
// one warp per row selects the two order statistics (radix select on order-preserving keys) and reduces the tail.
// `samples_out` (optional, [S][n]) receives the drawn samples so that tests can replay them through the oracle.
__global__ void __launch_bounds__(kRiskThreads) risk_mc_kernel(const float* __restrict__ mean,
                                                               const float* __restrict__ stdv, long long n,
                                                               const float* __restrict__ samples, int S, int S_pad,
                                                               int cpc, int metric, float q, uint32_t seed_lo,
                                                               uint32_t seed_hi, uint32_t ctr_lo, uint32_t ctr_hi,
                                                               float* __restrict__ risk, float* __restrict__ samples_out) {
  extern __shared__ __align__(16) float vals[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const long long cell0 = static_cast<long long>(blockIdx.x) * cpc;
  const int rs = S_pad + 1;  // row stride: +1 keeps the transposing fill free of bank conflicts
  if (samples != nullptr) {
    for (int idx = tid; idx < cpc * S_pad; idx += blockDim.x) {
      const int c = idx % cpc, s = idx / cpc;
      const long long cell = cell0 + c;
      float v = INFINITY;
      if (s < S && cell < n) v = samples[static_cast<size_t>(s) * n + cell];
      vals[c * rs + s] = v;
    }
  } else {
    const int quads = S_pad >> 2;  // S_pad is a multiple of 32
    const uint2 key = make_uint2(seed_lo, seed_hi);
    for (int idx = tid; idx < cpc * quads; idx += blockDim.x) {
      const int c = idx % cpc, j = idx / cpc;
      const long long cell = cell0 + c;
      float z[4] = {0.f, 0.f, 0.f, 0.f};
      float mu = 0.0f, sd = 0.0f;
      if (cell < n && 4 * j < S) {
        const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(cell), kStreamRisk | static_cast<uint32_t>(j),
                                                 ctr_lo, ctr_hi + static_cast<uint32_t>(cell >> 32)), key);
        const float2 a = box_muller(r.x, r.y), bq = box_muller(r.z, r.w);
        z[0] = a.x; z[1] = a.y; z[2] = bq.x; z[3] = bq.y;
        mu = mean[cell];
        sd = stdv[cell];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int s = 4 * j + u;
        float v = INFINITY;
        if (s < S && cell < n) {
          v = __fadd_rn(__fmul_rn(z[u], sd), mu);  // ATen normal: normal_(0,1).mul_(std).add_(mean)
          if (samples_out) samples_out[static_cast<size_t>(s) * n + cell] = v;
        }
        vals[c * rs + s] = v;
      }
    }
  }
  __syncthreads();
  const float rank = __fmul_rn(q, static_cast<float>(S - 1));  // ATen quantile: ranks = q * (n - 1) in the input dtype
  const float rlo_f = floorf(rank), rhi_f = ceilf(rank);
  const float wgt = __fsub_rn(rank, rlo_f);
  const unsigned int rlo = static_cast<unsigned int>(rlo_f), rhi = static_cast<unsigned int>(rhi_f);
  const int per_lane = S_pad >> 5;  // S_pad is a multiple of 32
  // One warp per cell.  Instead of sorting the row, the two order statistics torch.quantile interpolates between are
  // SELECTED: the samples are mapped to order-preserving unsigned keys and the rank-rlo key is found bit by bit from
  // the top (32 counting passes over the row, warp-reduced with REDUX); its successor needs one more pass.
  for (int c = warp; c < cpc; c += nwarps) {
    const long long cell = cell0 + c;
    if (cell >= n) break;
    unsigned int* keys = reinterpret_cast<unsigned int*>(vals + c * rs);
    for (int j = 0; j < per_lane; ++j) {  // float -> sortable key, in place (+inf pads sort last)
      const unsigned int bits = keys[j * 32 + lane];
      keys[j * 32 + lane] = bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
    }
    __syncwarp();
    unsigned int prefix = 0u, r = rlo;
    for (int bit = 31; bit >= 0; --bit) {
      unsigned int cnt0 = 0u;  // keys that agree with the prefix above `bit` and have this bit clear
      for (int j = 0; j < per_lane; ++j) cnt0 += (((keys[j * 32 + lane] ^ prefix) >> bit) == 0u) ? 1u : 0u;
      cnt0 = __reduce_add_sync(0xffffffffu, cnt0);
      if (r >= cnt0) {
        prefix |= 1u << bit;
        r -= cnt0;
      }
    }
    const unsigned int key_lo = prefix;
    unsigned int key_hi = key_lo;
    if (rhi != rlo) {  // the next order statistic: key_lo again if it has duplicates reaching rank rhi, else its successor
      unsigned int n_le = 0u, succ = 0xFFFFFFFFu;
      for (int j = 0; j < per_lane; ++j) {
        const unsigned int k = keys[j * 32 + lane];
        n_le += (k <= key_lo) ? 1u : 0u;
        if (k > key_lo) succ = min(succ, k);
      }
      n_le = __reduce_add_sync(0xffffffffu, n_le);
      succ = __reduce_min_sync(0xffffffffu, succ);
      key_hi = (n_le >= rhi + 1u) ? key_lo : succ;
    }
    auto key_to_float = [](unsigned int k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu)); };
    const float a = key_to_float(key_lo), bq = key_to_float(key_hi);
    // ATen lerp (vectorised CPU form): weight < 0.5 ? a + w (b - a) : b - (b - a)(1 - w), as one fused multiply-add
    const float diff = __fsub_rn(bq, a);
    const float var = (wgt < 0.5f) ? fmaf(wgt, diff, a) : fmaf(__fsub_rn(wgt, 1.0f), diff, bq);
    float out = var;
    if (metric == kRiskCvar) {
      float sum = 0.0f;
      int cnt = 0;
      for (int i = lane; i < S; i += 32) {
        const float v = key_to_float(keys[i]);
        if (v > var) {
          sum += v;
          ++cnt;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      }
      out = (cnt > 0) ? __fdiv_rn(sum, static_cast<float>(cnt)) : __int_as_float(0x7FC00000);
    }
    if (lane == 0) risk[cell] = out;
  }
}

// ------------------------------------------------------------------------------------------------ DWA
// dwa.py:160-184: window [max(u_min, a - a_lim dt), min(u_max, a + a_lim dt)] around the previous action a,
// torch.linspace over each axis (ATen: start + i step below the midpoint, end - (n-1-i) step above),
// torch.cartesian_prod(vs, omegas).  actions [nv*nw][2]; controls [nv*nw][T][2] = each action held over the horizon.
__global__ void __launch_bounds__(128) dwa_actions_kernel(const float* __restrict__ prev_action, Bounds b, float alim0,
                                                          float alim1, float dt, int nv, int nw, int T,
                                                          float* __restrict__ actions, float* __restrict__ controls) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nv * nw) return;
  const float p0 = prev_action[0], p1 = prev_action[1];
  const float v_lo = fmaxf(b.u_min0, __fsub_rn(p0, __fmul_rn(alim0, dt)));
  const float v_hi = fminf(b.u_max0, __fadd_rn(p0, __fmul_rn(alim0, dt)));
  const float w_lo = fmaxf(b.u_min1, __fsub_rn(p1, __fmul_rn(alim1, dt)));
  const float w_hi = fminf(b.u_max1, __fadd_rn(p1, __fmul_rn(alim1, dt)));
  auto lin = [](float lo, float hi, int n, int i) -> float {
    if (n == 1) return lo;
    const float step = __fdiv_rn(__fsub_rn(hi, lo), static_cast<float>(n - 1));
    return (i < n / 2) ? __fadd_rn(lo, __fmul_rn(step, static_cast<float>(i)))
                       : __fsub_rn(hi, __fmul_rn(step, static_cast<float>(n - i - 1)));
  };
  const float v = lin(v_lo, v_hi, nv, a / nw), w = lin(w_lo, w_hi, nw, a % nw);
  actions[2 * a] = v;
  actions[2 * a + 1] = w;
  float2* row = reinterpret_cast<float2*>(controls) + static_cast<size_t>(a) * T;
  for (int t = 0; t < T; ++t) row[t] = make_float2(v, w);
}

// dwa.py:225-228 + :260-285.  The reference selects the sub-goal from state_seq_batch[0, 0, :] AFTER the simulation:
// the raw (unclamped, unwrapped) successor that transit's in-place update left in slot 0 of the FIRST action's
// rollout, not the robot state.  Thread 0 recomputes that one step (traversability lookup in the engine's tau map,
// robot_model.py:75-88), then: among path points ahead of it (|atan2(d) - theta| < pi/2) and farther than the
// lookahead distance, the nearest one (first index whose distance equals that minimum); the last point if none.
// One CTA.
__global__ void __launch_bounds__(256) dwa_subgoal_kernel(GridGeom geom, int G, const float* __restrict__ tau,
                                                          int pitch, Bounds b, const float* __restrict__ actions,
                                                          const float* __restrict__ path, int N,
                                                          const float* __restrict__ state, float lookahead,
                                                          float* __restrict__ goal_out) {
  __shared__ float s_min[8];
  __shared__ int s_idx[8];
  __shared__ float s_best;
  __shared__ float s_ref[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    float x = state[0], y = state[1], th = state[2], xr, yr, thr;
    const float t0 = tau[cell_of(geom, G, pitch, x, y)];
    const float v0 = clampf(actions[0], b.u_min0, b.u_max0), v1 = clampf(actions[1], b.u_min1, b.u_max1);
    StepConsts c{};
    c.x_min = geom.x_min; c.y_min = geom.y_min; c.x_max = geom.x_max; c.y_max = geom.y_max; c.dt = b.dt;
    unicycle_step<false>(c, t0, v0, v1, x, y, th, xr, yr, thr);
    s_ref[0] = xr; s_ref[1] = yr; s_ref[2] = thr;
  }
  __syncthreads();
  const float sx = s_ref[0], sy = s_ref[1], sth = s_ref[2];
  float best = INFINITY;
  for (int i = tid; i < N; i += blockDim.x) {
    const float dx = __fsub_rn(path[2 * i], sx), dy = __fsub_rn(path[2 * i + 1], sy);
    const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    const float ang = __fsub_rn(atan2f(dy, dx), sth);
    if (fabsf(ang) < 1.57079637050628662f && d > lookahead) best = fminf(best, d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (lane == 0) s_min[warp] = best;
  __syncthreads();
  if (tid == 0) {
    float m = s_min[0];
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) m = fminf(m, s_min[w]);
    s_best = m;
  }
  __syncthreads();
  const float target = s_best;
  int first = 0x7FFFFFFF;
  if (target < INFINITY) {
    for (int i = tid; i < N; i += blockDim.x) {
      const float dx = __fsub_rn(path[2 * i], sx), dy = __fsub_rn(path[2 * i + 1], sy);
      const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      if (d == target) first = min(first, i);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
  if (lane == 0) s_idx[warp] = first;
  __syncthreads();
  if (tid == 0) {
    int f = s_idx[0];
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) f = min(f, s_idx[w]);
    const int pick = (target < INFINITY && f < N) ? f : N - 1;
    goal_out[0] = path[2 * pick];
    goal_out[1] = path[2 * pick + 1];
  }
}

// dwa.py:141-144: index of the minimum cost (first occurrence), that sample's action and recorded state sequence.
// One CTA.
__global__ void __launch_bounds__(256) argmin_gather_kernel(const float* __restrict__ costs, int K,
                                                            const float* __restrict__ actions,
                                                            const float* __restrict__ rec, int row_len,
                                                            float* __restrict__ action_out,
                                                            float* __restrict__ states_out, int* __restrict__ idx_out) {
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ int s_pick;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float bv = INFINITY;
  int bi = 0x7FFFFFFF;
  for (int i = tid; i < K; i += blockDim.x) {
    const float c = costs[i];
    if (c < bv || (c == bv && i < bi)) {
      bv = c;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov < bv || (ov == bv && oi < bi)) {
      bv = ov;
      bi = oi;
    }
  }
  if (lane == 0) {
    s_val[warp] = bv;
    s_idx[warp] = bi;
  }
  __syncthreads();
  if (tid == 0) {
    float v = s_val[0];
    int i0 = s_idx[0];
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w)
      if (s_val[w] < v || (s_val[w] == v && s_idx[w] < i0)) {
        v = s_val[w];
        i0 = s_idx[w];
      }
    if (i0 >= K) i0 = 0;  // all costs NaN/inf: fall back to the first sample
    s_pick = i0;
    if (idx_out) *idx_out = i0;
    action_out[0] = actions[2 * i0];
    action_out[1] = actions[2 * i0 + 1];
  }
  __syncthreads();
  const float* src = rec + static_cast<size_t>(s_pick) * row_len;
  for (int i = tid; i < row_len; i += blockDim.x) states_out[i] = src[i];
}

}  // namespace bnv
