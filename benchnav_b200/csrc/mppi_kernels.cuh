// sm_100a kernels of the MPPI control iteration (reference: src/planners/local_planners/mppi.py:130-240).
//
//   trav_map_kernel     risk map -> tau = 1 - clamp(risk,0,1), padded pitch            (traversability_model.py:71-72)
//   rollout_kernel      noise draw (Philox, interleaved with the rollout), clamp, T-step unicycle rollout, costs,
//                       per-CTA softmax partial; last CTA: grid merge, (sharded: exchange over NVLink), weights,
//                       optimal rollout (mppi.py:149-217).  Two variants (template flag kWide, see below).
//   normalize_weights_kernel  softmax weights of a launch whose grid was not co-resident
//   noise_kernel        the same Philox stream as a stand-alone kernel (tests / bnv_mppi_draw_noise); xi_kernel,
//                       slip_map_kernel, bump_iteration_kernel: stochastic-mode and graph-capture helpers
//   finalize_kernel     multi-GPU with a host-side exchange: merge gathered shard partials, weights, optimal rollout
//   top-n kernels       radix select + sort + gather (mppi.py:221-240); reroll_kernel for solvers without recorded states
//
// Work decomposition of rollout_kernel: one thread = one sample, state and running cost in registers; one warp = 32
// consecutive samples; warps never synchronise inside the T-loop; one CTA shares the traversability window staged by
// one 2-D TMA load.
//   Latency variant (kWide = false; every grid that fits the device in one co-resident wave, e.g. K = 16384): 4 warps
//   per CTA, one per SM sub-partition, each with whole-horizon slabs in shared memory -- noise (bulk-loaded from HBM when
//   injected, or produced in the loop and bulk-stored) and recorded states (bulk-stored).  At most one warp per
//   scheduler: the kernel is bound by the per-step dependency chain, so the loop body is branch-free, keeps every
//   invariant in registers, and fills the chain's stall slots with the next step pair's Philox draw.
//   Wide variant (kWide = true; single solvers that need several waves, e.g. K = 131072): 8 warps per CTA, two CTAs per
//   SM, <= 128 registers; recorded states and noise are staged in 16-step chunks and flushed with coalesced stores, the
//   weighted control sum re-reads the noise from L2, partials merge in two levels and the weights are normalised by
//   normalize_weights_kernel.  Bound by instruction issue, not by HBM.
//
// Two optional modes (template flags, separate instantiations so the single-solver path pays nothing):
//   kBatch  blockIdx.y = environment: E independent MPPI problems (own map, state, goal, mean sequence) in ONE
//           launch (BASELINE config 3); every per-environment buffer is the single-solver layout with a leading E.
//   kStoch  stochastic slip (BASELINE config 4): the window holds the cell's slip (mean, std); every lookup draws
//           tau = 1 - clamp(mean + std * xi, 0, 1) with a fresh xi ~ N(0,1) (traversability_model.py:65-69).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "mppi_math.cuh"
#include "ptx_sm100.cuh"

namespace bnv {

constexpr int kMaxWarps = 4;        // warps (x32 samples) per rollout CTA
constexpr int kFinalizeThreads = 128;
// Wide (throughput) variant of the rollout kernel, for grids that do not fit the device in one wave of the latency
// variant: 8 warps per CTA, two CTAs per SM, and instead of whole-horizon slabs each warp stages kChunkSteps steps of
// recorded states / drawn noise in shared memory and flushes them with coalesced stores.
constexpr int kWideWarps = 8;
constexpr int kChunkPairs = 8;
constexpr int kChunkSteps = 2 * kChunkPairs;
constexpr int kWideNzStride = 4 * kChunkPairs + 4;   // floats per sample row of the noise chunk: rows 16-byte aligned,
                                                     // per-lane 16-byte stores conflict-free (9 quad-banks apart)
constexpr int kWideRecStride = 3 * kChunkSteps + 1;  // floats per sample row of the recorded-state chunk (odd: conflict-free)
constexpr int kTwoLevelMin = 64;  // non-cooperative grids above this many CTAs merge their partials in two levels ...
constexpr int kMergeGroup = 32;   // ... in groups of this many consecutive CTAs

struct alignas(64) EngineParams {
  CUtensorMap tau_map;  // 2-D tiled descriptor over the padded tau map, box = patch_w x patch_h
  const float* tau;     // [E][G][pitch] traversability, or [E][G][pitch][2] slip (mean, std) in stochastic mode
  int G, pitch;         // pitch in cells
  GridGeom geom;
  Bounds bounds;
  float goal_x, goal_y, thr;
  float term_gx, term_gy;  // goal of the terminal cost: the goal itself, except under DWA's sub-goal (dwa.py:225-233)
  const float* goals;      // kBatch: [E][2] goal per environment (device); else optional [2] device override of goal_x/y
  int num_envs;            // E (1 unless kBatch)
  const float* xi_in;      // kStoch, injected: [E][Kl][2T+1] lookup normals (transit t, stage t interleaved; terminal last)
  const float* xi_opt_in;  // kStoch, injected: [E][T] lookup normals of the optimal rollout
  float lambda, inv_lambda, icov0, icov1;  // temperature, 1/sigma^2 (diag of mppi.py:95 inverse covariance)
  int lambda_pow2;                        // lambda is a power of two: -c * (1/lambda) == -c / lambda exactly
  float sigma0, sigma1;
  uint32_t seed_lo, seed_hi, iter_lo, iter_hi;  // Philox key and iteration counter
  int k_offset;                // global index of this shard's first sample
  int Kl, T;                   // shard-local samples, horizon
  int patch_w, patch_h, rho;   // staged window size (cells) and reach radius
  int use_patch;               // 0: window does not fit shared memory -> look up in the global map (L2)
  int world;                   // number of sample shards
  int record;                  // keep recorded states
  int rec_split;               // 0, or the (even) step Tc at which the recorded-state slab is flushed mid-loop: the slab
                               // then holds only max(Tc, T+1-Tc) slots per sample, which lets two CTAs share an SM when
                               // a grid would otherwise need a second wave (K = 32768 at T = 50)
  int noise_bulk_ok, rec_bulk_ok;  // pointers 16 B aligned -> bulk copies allowed
  const float* state;   // [3] device-resident current state, or
  float state_val[3];   // ... the same three floats passed by value in the launch packet (state_inline != 0)
  int state_inline;
  float* replay;        // [E][2T + 4] (lean solvers, record == 0): the mean sequence and the state this iteration
                        // started from, kept so that get_top_samples can re-roll the selected samples afterwards
  unsigned int dbg_flags;  // measurement aids (BNV_DEBUG_DISABLE >> 14): 2 = the optimal rollout draws its lookup
                           // normals in its own chain instead of reading the warp's pre-drawn ones
  int state_role;       // sharded solver driven from ONE rank's host (bnv_mppi_forward_host on the leader,
                        // bnv_mppi_forward_follow on the others): 1 = leader: broadcast state_val to every peer's state
                        // cell over NVLink at kernel start; 2 = follower: take the state from this rank's own cell
  const float* noise_in;  // injected noise [Kl][T][2] (kPhilox == false)
  float* noise_out;     // engine noise buffer [Kl][T][2], written when kPhilox
  float* u_prev;        // [T][2]   mean sequence (read at start, replaced by u* at the end)
  float* rec;           // [Kl][T+1][3]
  float* costs;         // [Kl]
  float* weights;       // [Kl]
  float* part_ms;       // [nCTA][2]   per-CTA (max score, sum exp)
  float* part_u;        // [nCTA][2T]  per-CTA sum exp * v
  float* shard_partial; // [2+2T]      (m, s, U) of this shard
  unsigned int* ticket_grp;    // [E][max_groups] arrival counters of the level-1 groups of a two-level merge
  float* part2_ms;             // [E][max_groups][2]   (max score, sum exp) per merged group
  float* part2_u;              // [E][max_groups][2T]  sum exp * v per merged group
  int max_groups;
  unsigned int* err_flag;      // device word: set when a peer exchange / hand-over wait timed out (result then invalid)
  unsigned int* ticket;  // [0] arrival counter of the last-CTA election, [1] "merge done" epoch flag (coop)
  float* stats;          // [2] (M, S) of the last merge, published to the waiting CTAs (coop)
  unsigned int epoch;    // unique per launch
  const unsigned long long* iter_dev;  // graph-capturable launches: the iteration counter lives in device memory (the
                                       // launch packet of a captured graph is frozen); then iteration = *iter_dev,
                                       // epoch = low word + 1, and bump_iteration_kernel advances it after the launch
  unsigned int* done_flag;  // optional, mapped HOST memory: [0] set to `epoch` once u_out / opt_rec are complete, [1] as
                            // soon as u_out alone is (signal_action), so that a
                            // host thread polling it sees the results without waiting for the kernel's tail
  // pre-launched iterations (bnv_mppi_prelaunch): the kernel is already resident when the state arrives.  It polls a
  // host-mapped slot {x, y, theta, sequence number} and every CTA follows one grid-wide decision (go / abort):
  const unsigned int* state_mailbox;   // mapped host memory, [2 slots][4 words]; null = state from P.state / P.state_val
  unsigned int mailbox_seq;            // sequence number this launch waits for (slot = seq & 1)
  unsigned int mailbox_timeout_us;     // give up (abort the launch) when no state arrives for this long
  unsigned int* prelaunch_decision;    // device word, zeroed before the launch: 0 undecided, 1 go, 2 abort
  unsigned int* abort_flag;            // mapped host word: set to `epoch` when the launch aborted
  int keep_mean;         // write u* back as the next call's mean sequence (mppi.py:217); 0 for DWA's constant actions
  int coop;              // grid is co-resident (cooperative launch): deferred slab stores, in-register weights
  float* const* peer_mbox;  // [world] device pointers to every rank's mailbox (peer memory over NVLink), or null
  float* peer_mbox_val[8];  // the same pointers by value for world <= 8 (one node): no table load on the exchange's path
  unsigned int xchg_seq;    // exchange sequence number (same on every rank), selects the mailbox parity
  int rank;
  float* u_out;         // [T][2]
  float* opt_rec;       // [T+1][3]
  long long* dbg_ts;    // optional [24] stamps (BNV_DEBUG_TS; null in production): clock64 phases of the last CTA in
                        // [0, 16), wall-clock (ns) stamps of the kernel start [16] and of column 0's exchange [8, 9, 11]
};

#define BNV_STAMP(i)                                                   \
  do {                                                                 \
    if (P.dbg_ts != nullptr && threadIdx.x == 0) P.dbg_ts[i] = clock64(); \
  } while (0)
#define BNV_STAMP_ANY(i, tid_)                                         \
  do {                                                                 \
    if (P.dbg_ts != nullptr && threadIdx.x == (tid_)) P.dbg_ts[i] = clock64(); \
  } while (0)

// Shared-memory carve-up of the rollout kernel (identical on host and device).
struct RolloutSmem {
  int off_patch, off_noise, off_rec, off_uprev, off_coef, off_e, off_warpu, off_red, off_merge, total;
};
constexpr int kMergeACap = 1280;    // fast grid merge: one copy PER WARP of the per-CTA rescale factors in shared memory
constexpr int kMergeGrpCap = 1024;  // ... and ngrp x 2T partial column sums
__host__ __device__ inline int rec_slab_slots(int T, int rec_split) {
  return rec_split > 0 ? (rec_split > T + 1 - rec_split ? rec_split : T + 1 - rec_split) : T + 1;
}
__host__ __device__ inline RolloutSmem rollout_smem_layout(int T, int warps, int patch_w, int patch_h, int use_patch,
                                                           int record, int cell_floats = 1, int rec_split = 0,
                                                           int wide = 0) {
  RolloutSmem s;
  int spb = warps * 32;
  int off = 128;  // [0,128): mbarriers (1 patch + kMaxWarps noise) and the last-CTA flag
  s.off_patch = off;
  off += use_patch ? ((patch_w * patch_h * cell_floats * 4 + 127) / 128) * 128 : 0;
  s.off_noise = off;
  off += (((wide ? spb * kWideNzStride : spb * 2 * T) * 4 + 127) / 128) * 128;
  s.off_rec = off;
  off += record ? (((wide ? spb * kWideRecStride : spb * 3 * rec_slab_slots(T, rec_split)) * 4 + 127) / 128) * 128 : 0;
  s.off_uprev = off;
  off += ((2 * T * 4 + 15) / 16) * 16;
  s.off_coef = off;
  off += 4 * T * 4;  // float4 per step
  s.off_e = off;
  off += spb * 4;
  s.off_warpu = off;
  off += (((warps > kMaxWarps ? warps : kMaxWarps) * 2 * T * 4 + 15) / 16) * 16;
  s.off_red = off;
  off += 64 * 4;
  s.off_merge = off;  // last CTA only: a_g [kMergeACap], group column sums [max(kMergeGrpCap, 2T)]
  off += (kMergeACap + (2 * T > kMergeGrpCap ? 2 * T : kMergeGrpCap)) * 4;
  s.total = off;
  return s;
}

#ifndef BNV_ROLLOUT_ONLY
// --------------------------------------------------------------------------------------------- tau map
// blockIdx.z = environment: `env_stride` elements between consecutive environments' risk maps (0 = shared map).
__global__ void trav_map_kernel(const float* __restrict__ risk, int risk_pitch, long long env_stride,
                                float* __restrict__ tau, int pitch, int G) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= pitch) return;
  const float* src = risk + static_cast<size_t>(blockIdx.z) * env_stride;
  float v = 0.0f;
  if (x < G) {
    float r = src[static_cast<size_t>(y) * risk_pitch + x];
    // torch.clamp propagates NaN; fminf/fmaxf would not
    float c = (r != r) ? r : fminf(fmaxf(r, 0.0f), 1.0f);
    v = __fsub_rn(1.0f, c);
  }
  tau[(static_cast<size_t>(blockIdx.z) * G + y) * pitch + x] = v;
}

// Stochastic mode: interleave the slip distribution's (mean, std) per cell, [E][G][pitch][2].
__global__ void slip_map_kernel(const float* __restrict__ mean, const float* __restrict__ stdv, int src_pitch,
                                long long env_stride, float2* __restrict__ out, int pitch, int G) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= pitch) return;
  const size_t off = static_cast<size_t>(blockIdx.z) * env_stride + static_cast<size_t>(y) * src_pitch + x;
  float2 v = make_float2(0.0f, 0.0f);
  if (x < G) v = make_float2(mean[off], stdv[off]);
  out[(static_cast<size_t>(blockIdx.z) * G + y) * pitch + x] = v;
}

// --------------------------------------------------------------------------------------------- noise
// Stand-alone draw of the engine's noise stream: one thread = one (sample, step pair) = noise[k][2p..2p+1][0..1].
// Bit-identical to what rollout_kernel<kPhilox> produces in its loop (same noise_pair()).  blockIdx.y = environment.
__global__ void __launch_bounds__(256) noise_kernel(float* __restrict__ noise, int Kl, int T, int k_offset,
                                                    uint32_t seed_lo, uint32_t seed_hi, uint32_t iter_lo,
                                                    uint32_t iter_hi, float sigma0, float sigma1) {
  int pairs = (T + 1) >> 1;
  long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(Kl) * pairs) return;
  int k = static_cast<int>(gid / pairs);
  int p = static_cast<int>(gid - static_cast<long long>(k) * pairs);
  const uint32_t env = blockIdx.y;
  const float4 n = noise_pair(static_cast<uint32_t>(k + k_offset), static_cast<uint32_t>(p), iter_lo,
                              iter_hi + (env << 16), make_uint2(seed_lo, seed_hi), sigma0, sigma1);
  float* dst = noise + ((static_cast<size_t>(env) * Kl + k) * T + 2 * p) * 2;
  *reinterpret_cast<float2*>(dst) = make_float2(n.x, n.y);
  if (2 * p + 1 < T) *reinterpret_cast<float2*>(dst + 2) = make_float2(n.z, n.w);
}

// Softmax weights of a grid that was not co-resident (launched right behind rollout_kernel): the rollout left
// exp(score_k - m_cta) in weights[k], the per-CTA maxima in part_ms and the merged (M, S) in stats (S = 1 when the 1/S
// is deferred to finalize_kernel).  weights[k] *= exp(m_cta - M) / S (mppi.py:193).  blockIdx.y = environment.
__global__ void __launch_bounds__(256) normalize_weights_kernel(float* __restrict__ weights,
                                                                const float* __restrict__ part_ms,
                                                                const float* __restrict__ stats, int Kl, int spb_shift,
                                                                int nblk, const unsigned int* __restrict__ decision) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Kl) return;
  if (decision != nullptr && *decision != 1u) return;  // the pre-launched rollout in front of this kernel aborted
  const size_t env = blockIdx.y;
  const float M = stats[2 * env], S = stats[2 * env + 1];
  const float m_cta = part_ms[(env * nblk + (k >> spb_shift)) * 2];
  float* w = weights + env * Kl + k;
  *w = *w * (__expf(m_cta - M) * __fdiv_rn(1.0f, S));
}

// Graph-capturable launches: advance the device-resident iteration counter after the rollout kernel.
__global__ void bump_iteration_kernel(unsigned long long* iter_dev) { *iter_dev += 1ull; }

// Stand-alone draw of the stochastic mode's lookup normals (same xi_quad() calls as the rollout kernel):
// xi [E][Kl][2T+1] = (transit 0, stage 0, transit 1, stage 1, ..., terminal), xi_opt [E][T] for the optimal rollout.
// One thread = one (sample, step pair); sample index Kl stands for the optimal rollout.
__global__ void __launch_bounds__(256) xi_kernel(float* __restrict__ xi, float* __restrict__ xi_opt, int Kl, int T,
                                                 int k_offset, uint32_t seed_lo, uint32_t seed_hi, uint32_t iter_lo,
                                                 uint32_t iter_hi) {
  const int pairs = (T + 1) >> 1;
  long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(Kl + 1) * pairs) return;
  const int k = static_cast<int>(gid / pairs);
  const int p = static_cast<int>(gid - static_cast<long long>(k) * pairs);
  const uint32_t env = blockIdx.y;
  const uint32_t ih = iter_hi + (env << 16);
  const uint2 key = make_uint2(seed_lo, seed_hi);
  if (k == Kl) {
    const float4 q = xi_quad(kOptimalSample, static_cast<uint32_t>(p), iter_lo, ih, key);
    float* dst = xi_opt + static_cast<size_t>(env) * T;
    dst[2 * p] = q.x;
    if (2 * p + 1 < T) dst[2 * p + 1] = q.z;
    return;
  }
  const uint32_t kg = static_cast<uint32_t>(k + k_offset);
  const float4 q = xi_quad(kg, static_cast<uint32_t>(p), iter_lo, ih, key);
  float* row = xi + (static_cast<size_t>(env) * Kl + k) * (2 * T + 1);
  row[4 * p] = q.x;
  row[4 * p + 1] = q.y;
  if (2 * p + 1 < T) {
    row[4 * p + 2] = q.z;
    row[4 * p + 3] = q.w;
  }
  if (p == 0) row[2 * T] = xi_quad(kg, kXiTerminalPair, iter_lo, ih, key).x;
}

#endif  // BNV_ROLLOUT_ONLY

// --------------------------------------------------------------------------------------------- helpers
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Window geometry for the current state: origin cell (ox, oy) and the clamp ranges.
struct WindowGeom {
  int ox, oy, lo_x, hi_x, lo_y, hi_y;
};
__device__ __forceinline__ WindowGeom window_for_state(const EngineParams& P, float sx, float sy) {
  WindowGeom w;
  if (!P.use_patch) {  // whole map, looked up in global memory
    w.ox = w.oy = w.lo_x = w.lo_y = 0;
    w.hi_x = w.hi_y = P.G - 1;
    return w;
  }
  int cx = min(max(cell_coord_rt(sx, P.geom.x_min, P.geom), 0), P.G - 1);
  int cy = min(max(cell_coord_rt(sy, P.geom.y_min, P.geom), 0), P.G - 1);
  // The innermost TMA coordinate must be a multiple of 16 bytes (4 cells) -- an unaligned x origin raises an
  // illegal-instruction fault on sm_100a -- so the origin is rounded down and patch_w carries 3 spare columns.
  // The box may overhang the map on the high side: overhanging cells are zero-filled and never indexed.
  w.ox = max(0, (cx - P.rho) & ~3);
  w.oy = max(0, cy - P.rho);
  w.lo_x = w.ox;
  w.hi_x = min(P.G - 1, w.ox + P.patch_w - 1);
  w.lo_y = w.oy;
  w.hi_y = min(P.G - 1, w.oy + P.patch_h - 1);
  return w;
}

// kCell = floats per window cell: 1 (traversability) or 2 (slip mean, std).  `tau_env` = this environment's map.
__device__ __forceinline__ StepConsts make_step_consts(const EngineParams& P, const WindowGeom& wg, const float* patch_s,
                                                       const float* tau_env, float gx, float gy, int cell_floats) {
  StepConsts c;
  c.x_min = P.geom.x_min; c.y_min = P.geom.y_min; c.x_max = P.geom.x_max; c.y_max = P.geom.y_max;
  c.res = P.geom.res; c.inv_res = P.geom.inv_res; c.dt = P.bounds.dt;
  c.gx = gx; c.gy = gy; c.thr = P.thr;
  c.u_min0 = P.bounds.u_min0; c.u_min1 = P.bounds.u_min1; c.u_max0 = P.bounds.u_max0; c.u_max1 = P.bounds.u_max1;
  c.lo_x = wg.lo_x; c.hi_x = wg.hi_x; c.lo_y = wg.lo_y; c.hi_y = wg.hi_y;
  const uint32_t esz = 4u * static_cast<uint32_t>(cell_floats);
  if (P.use_patch) {
    c.pitch = P.patch_w;
    c.win_addr = smem_u32(patch_s) - esz * static_cast<uint32_t>(wg.oy * P.patch_w + wg.ox);
    c.map = nullptr;
  } else {
    c.pitch = P.pitch;
    c.win_addr = 0u;
    c.map = tau_env;
  }
  c.finish(esz);
  return c;
}

// Score of a cost: x = -c / lambda (mppi.py:193), true division unless lambda is a power of two.
__device__ __forceinline__ float score_of(float cost, const EngineParams& P) {
  return P.lambda_pow2 ? __fmul_rn(-cost, P.inv_lambda) : __fdiv_rn(-cost, P.lambda);
}

// Where the lookup normals of the stochastic mode come from: injected (tests) or the engine's Philox stream.
struct XiSource {
  const float* row;  // injected: this sample's [2T+1] normals (or the optimal rollout's [T]); null = Philox
  uint32_t sample, iter_lo, iter_hi;
  uint2 key;
};

// Batch-1 optimal rollout (mppi.py:202-214) by one thread: first step with the general math, the rest fast.
// `uc_s` holds u* already clamped to the action bounds: transit re-clamps the action (robot_model.py:82-83) and
// u*, a rounded weighted sum, may leave the bounds by an ulp.  `out` = [T+1][3] recorded states.
// kStoch: every transit draws its own lookup normal (injected xi.row[t], or Philox sample id kOptimalSample).
template <bool kPatch, bool kPow2, bool kFastAngles, bool kStoch>
__device__ void optimal_rollout(int T, const StepConsts& c, const float* uc_s, float sx, float sy, float sth,
                                float* out, const XiSource& xi) {
  float x = sx, y = sy, th = sth, xr, yr, thr;
  float tau = 0.0f;
  float2 ms = make_float2(0.0f, 0.0f);
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  auto xi_at = [&](int t) -> float {
    if (xi.row != nullptr) return xi.row[t];  // injected (global) or pre-drawn by the warp (shared memory)
    if ((t & 1) == 0) q = xi_quad(xi.sample, static_cast<uint32_t>(t >> 1), xi.iter_lo, xi.iter_hi, xi.key);
    return (t & 1) ? q.z : q.x;
  };
  if (kStoch) {
    ms = lookup_slip<kPatch, kPow2, false>(c, x, y);
    tau = slip_to_trav(ms, xi_at(0));
  } else {
    tau = lookup_tau<kPatch, kPow2, false>(c, x, y);
  }
  const float2* u2 = reinterpret_cast<const float2*>(uc_s);
  float2 u = u2[0];
  unicycle_step<false>(c, tau, u.x, u.y, x, y, th, xr, yr, thr);
  out[0] = xr; out[1] = yr; out[2] = thr;
#pragma unroll 2
  for (int t = 1; t < T; ++t) {
    if (kStoch) {
      ms = lookup_slip<kPatch, kPow2, true>(c, x, y);
      tau = slip_to_trav(ms, xi_at(t));
    } else {
      tau = lookup_tau<kPatch, kPow2, true>(c, x, y);
    }
    u = u2[t];
    unicycle_step<kFastAngles>(c, tau, u.x, u.y, x, y, th, xr, yr, thr);
    out[3 * t + 0] = xr; out[3 * t + 1] = yr; out[3 * t + 2] = thr;
  }
  out[3 * T + 0] = x; out[3 * T + 1] = y; out[3 * T + 2] = th;
}

// Stochastic mode: the optimal rollout's T lookup normals do not depend on the rollout, so the warp draws them up
// front, one step pair per lane, into shared memory -- the serial thread then reads a float per step instead of running
// Philox4x32-10 + Box-Muller inside its dependency chain (measured: 490 -> ~150 cycles per step, 17k cycles per iteration
// of config 4).  Same xi_quad() calls as the in-chain draw: bit-identical.  All 32 lanes of the warp must call it.
__device__ __forceinline__ void predraw_optimal_xi(float* xi_s, int T, uint32_t iter_lo, uint32_t iter_hi, uint2 key,
                                                   int lane) {
  for (int p = lane; 2 * p < T; p += 32) {
    const float4 q = xi_quad(kOptimalSample, static_cast<uint32_t>(p), iter_lo, iter_hi, key);
    xi_s[2 * p] = q.x;
    if (2 * p + 1 < T) xi_s[2 * p + 1] = q.z;
  }
  __syncwarp();
}

// weights[k] *= scale(k): four samples per load, kBatch loads in flight per thread (the values were written by
// other SMs, so every load is an L2 round trip).  scale_of(g) gives the factor of CTA g's samples.
template <typename ScaleFn>
__device__ __forceinline__ void rescale_weights(float* weights, int Kl, int spb_shift, int wt, int wn, ScaleFn scale_of) {
  const int n4 = ((reinterpret_cast<uintptr_t>(weights) & 15u) == 0u) ? (Kl >> 2) : 0;  // float4 path needs alignment
  float4* w4 = reinterpret_cast<float4*>(weights);
  constexpr int kBatchLoads = 8;
  for (int base = 0; base < n4; base += wn * kBatchLoads) {
    float4 ev[kBatchLoads];
#pragma unroll
    for (int j = 0; j < kBatchLoads; ++j) {
      const int q = base + j * wn + wt;
      if (q < n4) ev[j] = __ldcg(w4 + q);
    }
#pragma unroll
    for (int j = 0; j < kBatchLoads; ++j) {
      const int q = base + j * wn + wt;
      if (q < n4) {
        const float a = scale_of((q << 2) >> spb_shift);  // spb is a multiple of 4: a float4 never straddles CTAs
        w4[q] = make_float4(ev[j].x * a, ev[j].y * a, ev[j].z * a, ev[j].w * a);
      }
    }
  }
  for (int kk = (n4 << 2) + wt; kk < Kl; kk += wn) weights[kk] = __ldcg(weights + kk) * scale_of(kk >> spb_shift);
}

// Last phase of a sharded iteration, run by ONE CTA of finalize_kernel once (M, S, U) over all shards are known:
// u* = U / S (mppi.py:196-199), next mean sequence (mppi.py:217), weights (mppi.py:193) and the batch-1
// optimal rollout (mppi.py:202-214).  `u_s` (shared, 2T floats) holds U on entry and u* on exit.
// `weights` holds exp(score - M_shard) per sample on entry; `w_scale` turns that into softmax weights.
template <bool kPatch, bool kPow2, bool kFastAngles>
__device__ void finish_iteration(const EngineParams& P, const StepConsts& c, float S, float w_scale, float* u_s,
                                 float sx, float sy, float sth) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int T = P.T;
  float* uc_s = u_s + ((2 * T + 3) & ~3);  // clamped copy for the optimal rollout
  for (int i = tid; i < 2 * T; i += nthr) {
    float u = __fdiv_rn(u_s[i], S);
    u_s[i] = u;
    uc_s[i] = clampf(u, (i & 1) ? c.u_min1 : c.u_min0, (i & 1) ? c.u_max1 : c.u_max0);
    P.u_out[i] = u;
    P.u_prev[i] = u;
  }
  __syncthreads();
  if (tid == 0) {
    XiSource none{};
    optimal_rollout<kPatch, kPow2, kFastAngles, false>(T, c, uc_s, sx, sy, sth, P.opt_rec, none);
  }
  // With more than one warp the serial optimal rollout keeps warp 0 busy and the other warps rescale underneath it.
  const bool split = nthr > 32;
  if (split && tid < 32) return;
  if (!split) __syncwarp();
  rescale_weights(P.weights, P.Kl, 0, split ? tid - 32 : tid, split ? nthr - 32 : nthr,
                  [w_scale](int) { return w_scale; });
}

// --------------------------------------------------------------------------------------------- rollout
// Wide variant: copy a block of 32 rows x n words from a shared-memory chunk slab (row stride `src_stride` words) to
// global rows (row stride `dst_stride` words), all 32 lanes on consecutive words of the flat (row, word) index space,
// so that every store instruction covers whole 128-byte runs except where it crosses a row boundary.  Generic in n:
// used for the last, partial chunk (the full chunk has its own specialisation below).
__device__ __forceinline__ void copy_rows_flat(float* __restrict__ dst, int dst_stride, const float* __restrict__ src,
                                               int src_stride, int n, int lane) {
  const uint32_t magic = 0xFFFFFFFFu / static_cast<uint32_t>(n) + 1u;  // ceil(2^32 / n): row = umulhi(idx, magic), idx < 2^16
  const int total = 32 * n;
#pragma unroll 1
  for (int base = 0; base < total; base += 256) {  // eight 128-byte passes in flight
    float v[8];
    int off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = base + 32 * j + lane;
      const int row = static_cast<int>(__umulhi(static_cast<uint32_t>(idx), magic));
      const int col = idx - row * n;
      off[j] = idx < total ? row * dst_stride + col : -1;
      v[j] = idx < total ? src[row * src_stride + col] : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (off[j] >= 0) dst[off[j]] = v[j];
  }
}

// Coalesced warp copy of recorded-state slots [t0, t0 + nt) of `rows` samples from the slab (row stride `slab_slots`
// slots, slot t0 at the row start) to HBM rows of T+1 slots: used when the slab is flushed in two halves (a half
// row is not a 16-byte multiple, so it cannot go out as one bulk copy).  All 32 lanes participate.  A full warp goes
// through the flat copy (eight 128-byte passes in flight, lanes run across row boundaries) instead of three dependent
// load/store passes per row.
__device__ __forceinline__ void copy_rec_slots(float* rec_g, const float* rec_w, int rows, int T, int slab_slots, int t0,
                                               int nt, int lane) {
  const int n = 3 * nt;
  if (rows == 32 && n < 2048) {
    copy_rows_flat(rec_g + 3 * t0, 3 * (T + 1), rec_w, 3 * slab_slots, n, lane);
    return;
  }
  for (int r = 0; r < rows; ++r) {
    float* dst = rec_g + static_cast<size_t>(r) * 3 * (T + 1) + 3 * t0;
    const float* src = rec_w + r * 3 * slab_slots;
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
  }
}

// The full chunk of recorded states (32 rows x 48 words), specialised: 48 words per row and 32 lanes per pass repeat
// every 3 passes = 2 rows, so each lane needs only three (shared, global) word offsets, computed once; per row pair
// the copy is 3 LDS + 3 STG with immediate shared-memory offsets (the generic flat copy spends ~9 instructions per
// pass on index arithmetic -- measured 21 % of the wide variant's instructions before this).
struct RecFlushLane {
  int so[3], go[3];  // word offsets within a row pair: shared (row stride kWideRecStride) / global (row stride 3 (T+1))
};
__device__ __forceinline__ RecFlushLane rec_flush_lane(int T, int lane) {
  RecFlushLane f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int w = 32 * j + lane;  // 0 .. 95 = two rows of 48 words
    const int second = w >= 3 * kChunkSteps ? 1 : 0;
    const int col = w - second * 3 * kChunkSteps;
    f.so[j] = second * kWideRecStride + col;
    f.go[j] = second * 3 * (T + 1) + col;
  }
  return f;
}
__device__ __forceinline__ void flush_rec_chunk(float* __restrict__ dst, int T, const float* __restrict__ src,
                                                const RecFlushLane& f) {
  const int pair_stride = 2 * 3 * (T + 1);
#pragma unroll
  for (int q0 = 0; q0 < 16; q0 += 4) {  // four row pairs = twelve loads in flight
    float v[4][3];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 3; ++j) v[q][j] = src[(q0 + q) * 2 * kWideRecStride + f.so[j]];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 3; ++j) dst[(q0 + q) * pair_stride + f.go[j]] = v[q][j];
  }
}

// Slabs -> HBM: recorded states (and the drawn noise), one bulk store per warp slab; asynchronous, the caller
// waits for the shared-memory reads (bulk_wait_read_all) before the CTA exits.  rec_env / noise_env = this
// environment's [Kl][T+1][3] / [Kl][T][2] arrays.
template <bool kRecord, bool kPhilox>
__device__ __forceinline__ void store_slabs(const EngineParams& P, float* rec_env, float* noise_env, float* rec_s,
                                            float* nz_w, int warp, int lane, int warp_first, int warp_rows,
                                            uint32_t nz_bytes) {
  if (warp_rows <= 0 || !(kRecord || kPhilox)) return;
  const int T = P.T;
  const int slots = rec_slab_slots(T, P.rec_split);
  float* rec_g = rec_env + static_cast<size_t>(warp_first) * 3 * (T + 1);
  const uint32_t rec_bytes = static_cast<uint32_t>(warp_rows) * 3u * (T + 1) * 4u;
  float* rec_w = rec_s + warp * 32 * 3 * slots;
  float* nz_g = noise_env + static_cast<size_t>(warp_first) * 2 * T;
  const bool rec_bulk = kRecord && P.rec_split == 0 && P.rec_bulk_ok && (rec_bytes & 15u) == 0u &&
                        (reinterpret_cast<uintptr_t>(rec_g) & 15u) == 0u;
  const bool out_bulk = kPhilox && (nz_bytes & 15u) == 0u && (reinterpret_cast<uintptr_t>(nz_g) & 15u) == 0u;
  fence_proxy_async_smem();
  __syncwarp();
  if (elect_one()) {
    if (rec_bulk) bulk_store_s2g(rec_g, rec_w, rec_bytes);
    if (out_bulk) bulk_store_s2g(nz_g, nz_w, nz_bytes);
    bulk_commit();
  }
  if (kRecord && !rec_bulk) {
    if (P.rec_split > 0) copy_rec_slots(rec_g, rec_w, warp_rows, T, slots, P.rec_split, T + 1 - P.rec_split, lane);
    else
      for (int i = lane; i < warp_rows * 3 * (T + 1); i += 32) rec_g[i] = rec_w[i];
  }
  if (kPhilox && !out_bulk)
    for (int i = lane; i < warp_rows * 2 * T; i += 32) nz_g[i] = nz_w[i];
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int atom_add_acq_rel_gpu(unsigned int* p, unsigned int v) {
  unsigned int old;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Self-validating 8-byte words {value, tag} ("LL" protocol): one aligned 8-byte store is a single memory transaction,
// so a reader that sees the expected tag also sees the value -- no fence, no separate flag.  volatile = relaxed at
// system scope: the same two functions serve hand-overs inside the GPU and stores into a peer GPU's memory over NVLink.
__device__ __forceinline__ void st_ll(uint2* p, float v, uint32_t tag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const uint2* p) {
  uint2 w;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(p) : "memory");
  return w;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kSpinLimitNs = 2000000000ull;  // 2 s: a peer that never answers must not hang the device
// Called every 1024 polls of a wait loop: starts the clock on the first call, true once the limit has passed.
__device__ __forceinline__ bool spin_expired(unsigned long long& t0) {
  const unsigned long long now = globaltimer_ns();
  if (t0 == 0ull) {
    t0 = now;
    return false;
  }
  return now - t0 > kSpinLimitNs;
}

// Wait for the LL word at `p` to carry `tag`; returns its value.  On a time-out the solver's error word is raised and
// 0 is returned (the iteration's result is then invalid; bnv_mppi_check reports it).
__device__ __forceinline__ float wait_ll(const uint2* p, uint32_t tag, unsigned int* err_flag) {
  uint2 w = ld_ll(p);
  if (w.y == tag) return __uint_as_float(w.x);
  unsigned long long t0 = 0ull;  // (read lazily: %globaltimer is slow, and normally nobody waits long enough to need it)
  for (unsigned int spins = 1;; ++spins) {
    w = ld_ll(p);
    if (w.y == tag) return __uint_as_float(w.x);
    if ((spins & 1023u) == 0u && spin_expired(t0)) {
      if (err_flag != nullptr) atomicExch(err_flag, 1u);
      return 0.0f;
    }
  }
}

// Sharded softmax, one column of the exchange, executed by ONE thread: this rank's un-normalised column
// sum U[c] (relative to its shard maximum M, with the shard's sum S) goes into the cell (rank, c) of every rank's
// mailbox as three LL words over NVLink peer memory; then the thread collects the W cells of column c from its own
// mailbox and merges them in rank order -- identical arithmetic on every rank, so every rank holds the same u*.
// Mailboxes are double-buffered by the parity of the exchange sequence number (= the tag): a rank can be at most one
// iteration ahead of the slowest peer.  Mailbox layout: [2 parities][W ranks][2T columns][3] uint2.
struct ColumnMerge {
  float U, M, S;
};
__device__ __forceinline__ const uint2* mailbox_cell(const EngineParams& P, int holder, int src_rank, int c, int ncol) {
  const size_t idx = ((static_cast<size_t>(P.xchg_seq & 1u) * P.world + src_rank) * ncol + c) * 3;
  const float* base = P.world <= 8 ? P.peer_mbox_val[holder] : P.peer_mbox[holder];
  return reinterpret_cast<const uint2*>(base) + idx;
}
// Collect the cells (src ranks r0 .. r0 + n - 1, n <= 8) of column c from this rank's own mailbox: all outstanding
// words are requested before the first is examined, so a round costs one L2 round trip; rounds repeat until every word
// carries the tag (or the time-out raises the error word).
__device__ __forceinline__ void poll_cells(const EngineParams& P, int c, int ncol, int r0, int n, float (&U)[8],
                                           float (&M)[8], float (&S)[8]) {
  const uint32_t tag = P.xchg_seq;
  unsigned pend = 0u;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    U[j] = 0.0f;
    M[j] = -FLT_MAX;
    S[j] = 0.0f;
    if (j < n) pend |= 7u << (3 * j);
  }
  const uint2* cell0 = mailbox_cell(P, P.rank, r0, c, ncol);
  const size_t rank_stride = static_cast<size_t>(ncol) * 3;
  unsigned long long t0 = 0ull;  // (read lazily: %globaltimer is slow, and normally nobody waits long enough to need it)
  for (unsigned int spins = 1; pend != 0u; ++spins) {
    uint2 w[8][3];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        w[j][q] = make_uint2(0u, 0u);
        if (pend & (1u << (3 * j + q))) w[j][q] = ld_ll(cell0 + j * rank_stride + q);
      }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int q = 0; q < 3; ++q)
        if ((pend & (1u << (3 * j + q))) && w[j][q].y == tag) {
          const float v = __uint_as_float(w[j][q].x);
          if (q == 0) U[j] = v;
          if (q == 1) M[j] = v;
          if (q == 2) S[j] = v;
          pend &= ~(1u << (3 * j + q));
        }
    if ((spins & 1023u) == 0u && spin_expired(t0)) {
      if (P.err_flag != nullptr) atomicExch(P.err_flag, 1u);
      pend = 0u;
    }
  }
}
// The leader's state for followers: three LL words {x | y | theta, tag} behind the column cells of the mailbox,
// double-buffered by the same parity.
__device__ __forceinline__ const uint2* mailbox_state_cell(const EngineParams& P, int holder, int ncol) {
  const float* base = P.world <= 8 ? P.peer_mbox_val[holder] : P.peer_mbox[holder];
  return reinterpret_cast<const uint2*>(base) + static_cast<size_t>(2) * P.world * ncol * 3 + (P.xchg_seq & 1u) * 3;
}
__device__ __forceinline__ ColumnMerge merge_column_cells(const EngineParams& P, int c, int ncol) {
  const int W = P.world;
  float U[8], M[8], S[8];
  float Mg = -FLT_MAX;
  for (int r0 = 0; r0 < W; r0 += 8) {  // (one group on a single node: W <= 8)
    poll_cells(P, c, ncol, r0, min(8, W - r0), U, M, S);
#pragma unroll
    for (int j = 0; j < 8; ++j) Mg = fmaxf(Mg, M[j]);
  }
  ColumnMerge out{0.0f, Mg, 0.0f};
  for (int r0 = 0; r0 < W; r0 += 8) {  // rank order: the same sums on every rank
    if (W > 8) poll_cells(P, c, ncol, r0, min(8, W - r0), U, M, S);  // (all valid by now: one round)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (r0 + j < W) {
        const float a = __expf(M[j] - Mg);
        out.S = fmaf(a, S[j], out.S);
        out.U = fmaf(a, U[j], out.U);
      }
    }
  }
  return out;
}
__device__ __forceinline__ ColumnMerge exchange_column(const EngineParams& P, int c, int ncol, float U, float M, float S) {
  const uint32_t tag = P.xchg_seq;
  const bool stamp = P.dbg_ts != nullptr && c == 0;  // BNV_DEBUG_TS: wall-clock (ns) stamps of column 0's exchange
  if (stamp) P.dbg_ts[8] = static_cast<long long>(globaltimer_ns());
  for (int r = 0; r < P.world; ++r) {
    uint2* cell = const_cast<uint2*>(mailbox_cell(P, r, P.rank, c, ncol));
    st_ll(cell, U, tag);
    st_ll(cell + 1, M, tag);
    st_ll(cell + 2, S, tag);
  }
  if (stamp) P.dbg_ts[9] = static_cast<long long>(globaltimer_ns());
  const ColumnMerge out = merge_column_cells(P, c, ncol);
  if (stamp) P.dbg_ts[11] = static_cast<long long>(globaltimer_ns());
  return out;
}

// Results complete (u_out and opt_rec written by this CTA, ordered before this thread by the CTA barrier): publish
// them system-wide and raise the host-visible completion word.
__device__ __forceinline__ void signal_done(const EngineParams& P) {
  if (P.done_flag != nullptr) {
    __threadfence_system();
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(P.done_flag), "r"(P.epoch) : "memory");
  }
}

// First stage of the host-visible completion: u* is written (by other threads of this CTA, ordered before this thread by
// the CTA / named barrier).  Raised by a thread that is NOT on the path to the serial optimal rollout, so a host that
// only needs the controls (the next environment step) can go on ~3.7 us before the optimal state sequence is done.
__device__ __forceinline__ void signal_action(const EngineParams& P) {
  if (P.done_flag != nullptr) {
    __threadfence_system();
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(P.done_flag + 1), "r"(P.epoch) : "memory");
  }
}

// One rollout step of one sample (mppi.py:152-165, :174-182): control from mean + noise, unicycle step, recorded
// state, shared lookup for the stage cost and the next step, cost accumulation.
struct SampleState {
  float x, y, th, tau, stage_sum, act0, act1;
  float2 ms;  // kStoch: (mean, std) of the slip distribution in the current cell
};

// ucf_s[t] = (u_prev[t][0], u_prev[t][1], u_prev[t][0] / sigma0^2, u_prev[t][1] / sigma1^2): one 16-byte load per step.
// The clamped control v is not kept: the weighted-sum pass recomputes it from the noise slab with the same two ops.
// kStoch: xi_tr / xi_st = lookup normals of this step's transit and of the stage cost of its recorded state; the two
// lookups hit the same cell (the index clamp makes cell(raw successor) == cell(clamped successor)), so the (mean, std)
// fetch is shared and only the draw differs.
template <bool kPatch, bool kPow2, bool kRecord, bool kFastStep, bool kStoch>
__device__ __forceinline__ void sample_step(SampleState& s, const StepConsts& C, int t, float nx, float ny, float xi_tr,
                                            float xi_st, const float4* ucf_s, float* rrow) {
  const float4 uc = ucf_s[t];
  const float v0 = clampf(__fadd_rn(uc.x, nx), C.u_min0, C.u_max0);  // mppi.py:152-157
  const float v1 = clampf(__fadd_rn(uc.y, ny), C.u_min1, C.u_max1);
  float xr, yr, thr;
  const float tau_dyn = kStoch ? slip_to_trav(s.ms, xi_tr) : s.tau;
  unicycle_step<kFastStep>(C, tau_dyn, v0, v1, s.x, s.y, s.th, xr, yr, thr);
  if (kRecord) {
    rrow[3 * t + 0] = xr;
    rrow[3 * t + 1] = yr;
    rrow[3 * t + 2] = thr;
  }
  // one lookup serves the stage cost of the recorded (raw) position and the next dynamics step
  float tau_cost;
  if (kStoch) {
    s.ms = lookup_slip<kPatch, kPow2, true>(C, s.x, s.y);
    tau_cost = slip_to_trav(s.ms, xi_st);
  } else {
    s.tau = lookup_tau<kPatch, kPow2, true>(C, s.x, s.y);
    tau_cost = s.tau;
  }
  s.stage_sum = __fadd_rn(s.stage_sum, goal_and_stuck_cost(C, xr, yr, tau_cost));
  s.act0 = fmaf(uc.z, v0, s.act0);  // u_prev[t]^T Sigma^-1 v (mppi.py:178-182), one running sum per control dim
  s.act1 = fmaf(uc.w, v1, s.act1);
}

// kPatch: traversability window staged in shared memory by TMA (else looked up in the global map);
// kPow2: resolution is a power of two (multiply instead of divide in the cell index);
// kRecord: keep every sample's recorded states; kFastAngles: dt * max|omega| < pi (branch-free steps 1..T-1);
// kPhilox: draw the noise in the loop (else it is injected and bulk-loaded from HBM);
// kStoch: stochastic-slip lookups; kBatch: blockIdx.y = environment;
// kWide: throughput variant -- 8 warps per CTA, 2 CTAs per SM (<= 128 registers), chunked staging of the recorded
// states and the drawn noise, weighted sum re-reads the noise from L2/HBM; never launched cooperatively.
template <bool kPatch, bool kPow2, bool kRecord, bool kFastAngles, bool kPhilox, bool kStoch, bool kBatch,
          bool kWide = false>
__global__ void __launch_bounds__((kWide ? kWideWarps : kMaxWarps) * 32, kWide ? 2 : 1)
    rollout_kernel(const __grid_constant__ EngineParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const long long t_start = clock64();
  if (P.dbg_ts != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    P.dbg_ts[16] = static_cast<long long>(globaltimer_ns());  // BNV_DEBUG_TS: wall clock (ns) at the start of CTA 0
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = P.T;
  const int nwarps = blockDim.x >> 5;
  const int spb = nwarps * 32;
  constexpr int kCell = kStoch ? 2 : 1;
  const RolloutSmem L = rollout_smem_layout(T, nwarps, P.patch_w, P.patch_h, kPatch, kRecord, kCell, P.rec_split, kWide);
  const int rec_slots = rec_slab_slots(T, P.rec_split);  // slots per sample in the recorded-state slab
  uint64_t* bar_patch = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar_noise = reinterpret_cast<uint64_t*>(smem) + 1;  // [kMaxWarps]
  int* last_flag = reinterpret_cast<int*>(smem + 64);
  float* patch_s = reinterpret_cast<float*>(smem + L.off_patch);
  float* noise_s = reinterpret_cast<float*>(smem + L.off_noise);
  float* rec_s = reinterpret_cast<float*>(smem + L.off_rec);
  float* uprev_s = reinterpret_cast<float*>(smem + L.off_uprev);
  float* coef_s = reinterpret_cast<float*>(smem + L.off_coef);
  float* e_s = reinterpret_cast<float*>(smem + L.off_e);
  float* warpu_s = reinterpret_cast<float*>(smem + L.off_warpu);
  float* red_s = reinterpret_cast<float*>(smem + L.off_red);

  // ---- this CTA's environment: every per-solver array carries a leading E in batch mode
  const int env = kBatch ? static_cast<int>(blockIdx.y) : 0;
  const size_t eK = static_cast<size_t>(env) * P.Kl;                 // first sample row of the environment
  const size_t eB = static_cast<size_t>(env) * gridDim.x;            // first per-CTA partial of the environment
  float* const u_prev_e = P.u_prev + static_cast<size_t>(env) * 2 * T;
  float* const costs_e = P.costs + eK;
  float* const weights_e = P.weights + eK;
  float* const rec_e = kRecord ? P.rec + eK * 3 * (T + 1) : nullptr;
  float* const noise_out_e = P.noise_out + eK * 2 * T;
  float* const part_ms_e = P.part_ms + eB * 2;
  float* const part_u_e = P.part_u + eB * 2 * T;
  unsigned int* const ticket_e = P.ticket + 2 * env;  // [0] arrival counter, [1] "merge done" epoch flag
  float* const stats_e = P.stats + 2 * env;
  // iteration counter and launch epoch: by value, or from device memory when the launch is replayed from a CUDA graph
  uint32_t iter_lo = P.iter_lo, iter_hi = P.iter_hi, epoch = P.epoch;
  if (P.iter_dev != nullptr) {
    const unsigned long long it = __ldcg(P.iter_dev);
    iter_lo = static_cast<uint32_t>(it);
    iter_hi = static_cast<uint32_t>(it >> 32);
    // epoch = it mod (2^32 - 1) + 1: never 0, and different for consecutive iterations for ever (2^32 = 1 mod 2^32 - 1)
    unsigned long long f = (it >> 32) + (it & 0xFFFFFFFFull);
    f = (f >> 32) + (f & 0xFFFFFFFFull);
    if (f >= 0xFFFFFFFFull) f -= 0xFFFFFFFFull;
    epoch = static_cast<uint32_t>(f) + 1u;
  }
  const uint32_t iter_hi_e = iter_hi + (static_cast<uint32_t>(env) << 16);  // Philox counter word 3 carries the env

  const int cta_first = blockIdx.x * spb;
  const int warp_first = cta_first + warp * 32;
  const int warp_rows = max(0, min(32, P.Kl - warp_first));
  const int k = warp_first + lane;
  const bool valid = lane < warp_rows;

  // the state (a cold load when it lives in HBM) is requested first: its latency runs under the barrier set-up
  float sx, sy, sth;
  if (P.state_mailbox != nullptr) {
    // Pre-launched iteration.  ONE thread of the grid (CTA 0) polls the host-mapped slot over PCIe until it carries
    // this launch's sequence number (measured: 128 CTAs polling host memory serialise at ~1 us per read), then
    // publishes the state and the decision "go" in device memory; a cancel mark or a timeout publishes "abort".
    // Every other CTA polls the decision word in L2.
    unsigned int* box_s = reinterpret_cast<unsigned int*>(smem + 96);
    if (tid == 0) {
      unsigned int decision = 0u, w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
      unsigned int* dec = P.prelaunch_decision;  // [0] decision, [4..6] state words
      if (blockIdx.x == 0 && blockIdx.y == 0) {
        const unsigned int* slot = P.state_mailbox + 4u * (P.mailbox_seq & 1u);
        unsigned long long t0, now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        const unsigned long long limit = static_cast<unsigned long long>(P.mailbox_timeout_us) * 1000ull;
        for (;;) {
          asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                       : "l"(slot)
                       : "memory");
          if (w3 == P.mailbox_seq) {
            decision = 1u;
            break;
          }
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (w3 == 0xFFFFFFFFu || now - t0 > limit) {
            decision = 2u;
            break;
          }
        }
        if (decision == 1u) {
          dec[4] = w0;
          dec[5] = w1;
          dec[6] = w2;
        } else {
          asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(P.abort_flag), "r"(P.epoch) : "memory");
        }
        st_release_gpu(dec, decision);
      } else {
        while ((decision = ld_acquire_gpu(dec)) == 0u) __nanosleep(100);
        if (decision == 1u) {
          w0 = __ldcg(dec + 4);
          w1 = __ldcg(dec + 5);
          w2 = __ldcg(dec + 6);
        }
      }
      box_s[0] = w0;
      box_s[1] = w1;
      box_s[2] = w2;
      box_s[3] = (decision == 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (box_s[3] == 0u) return;  // aborted launch: nothing was touched
    sx = __uint_as_float(box_s[0]);
    sy = __uint_as_float(box_s[1]);
    sth = __uint_as_float(box_s[2]);
  } else {
    if (!kBatch && P.state_role == 2) {
      // follower of a host-driven sharded solver: the leader's kernel stores the state into this rank's mailbox over
      // NVLink (three LL words); thread 0 of every CTA polls the (local) cell
      float* box_f = reinterpret_cast<float*>(smem + 96);
      if (tid == 0) {
        const uint2* cell = mailbox_state_cell(P, P.rank, 2 * T);
        box_f[0] = wait_ll(cell, P.xchg_seq, P.err_flag);
        box_f[1] = wait_ll(cell + 1, P.xchg_seq, P.err_flag);
        box_f[2] = wait_ll(cell + 2, P.xchg_seq, P.err_flag);
      }
      __syncthreads();
      sx = box_f[0];
      sy = box_f[1];
      sth = box_f[2];
    } else {
      const float* state_e = P.state + 3 * env;
      sx = P.state_inline ? P.state_val[0] : __ldg(state_e);
      sy = P.state_inline ? P.state_val[1] : __ldg(state_e + 1);
      sth = P.state_inline ? P.state_val[2] : __ldg(state_e + 2);
      if (!kBatch && P.state_role == 1 && blockIdx.x == 0 && tid < P.world && tid != P.rank) {
        uint2* cell = const_cast<uint2*>(mailbox_state_cell(P, tid, 2 * T));  // leader: one thread per peer
        st_ll(cell, sx, P.xchg_seq);
        st_ll(cell + 1, sy, P.xchg_seq);
        st_ll(cell + 2, sth, P.xchg_seq);
      }
    }
  }
  if (tid == 0) {
    mbar_init(bar_patch, 1);
    for (int w = 0; w < kMaxWarps; ++w) mbar_init(bar_noise + w, 1);
    fence_mbar_init();
    if (kPatch) prefetch_tensormap(&P.tau_map);
  }
  __syncthreads();

  // ---- injected noise: stage this warp's slab, rows [warp_first, warp_first+warp_rows) x 2T floats, contiguous in HBM
  float* nz_w = noise_s + warp * 32 * (kWide ? kWideNzStride : 2 * T);
  const uint32_t nz_bytes = static_cast<uint32_t>(warp_rows) * 2u * T * 4u;
  bool nz_bulk = false;
  if (!kPhilox && !kWide) {  // (wide: every thread reads its own injected row from global memory in the loop)
    const float* nz_g = P.noise_in + (eK + warp_first) * 2 * T;
    nz_bulk = P.noise_bulk_ok && ((nz_bytes & 15u) == 0u) && warp_rows > 0 &&
              (reinterpret_cast<uintptr_t>(nz_g) & 15u) == 0u;
    if (nz_bulk) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_noise + warp, nz_bytes);
        bulk_load_g2s(nz_w, nz_g, nz_bytes, bar_noise + warp);
      }
    } else {
      for (int i = lane; i < warp_rows * 2 * T; i += 32) nz_w[i] = nz_g[i];
    }
  }
  // noise rows of a ragged warp that carry no sample: keep them finite (the weighted-sum pass multiplies the
  // controls rebuilt from them by 0)
  if (!kWide)
    for (int i = warp_rows * 2 * T + lane; i < 32 * 2 * T; i += 32) nz_w[i] = 0.0f;

  // ---- window geometry, traversability window via TMA
  const WindowGeom wg = window_for_state(P, sx, sy);
  if (kPatch) {
    if (warp == 0) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_patch, static_cast<uint32_t>(P.patch_w * P.patch_h * 4 * kCell));
        tma_load_2d(patch_s, &P.tau_map, wg.ox * kCell, wg.oy + env * P.G, bar_patch);
      }
    }
  }
  const float* tau_e = P.tau + static_cast<size_t>(env) * P.G * P.pitch * kCell;
  // goal: per environment (batch), a device-resident override (DWA's sub-goal, selected on the device), or by value
  const float gx = kBatch ? __ldg(P.goals + 2 * env) : (P.goals != nullptr ? __ldg(P.goals) : P.goal_x);
  const float gy = kBatch ? __ldg(P.goals + 2 * env + 1) : (P.goals != nullptr ? __ldg(P.goals + 1) : P.goal_y);
  const float term_gx = kBatch ? gx : P.term_gx, term_gy = kBatch ? gy : P.term_gy;
  StepConsts C = make_step_consts(P, wg, patch_s, tau_e, gx, gy, kCell);

  // first step pair of the Philox stream: drawn while the TMA window and the mean sequence are in flight
  const uint32_t kg = static_cast<uint32_t>(k + P.k_offset);
  const uint2 key = make_uint2(P.seed_lo, P.seed_hi);
  float sig0 = P.sigma0, sig1 = P.sigma1;
  float4 nz_cur = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 xi_cur = make_float4(0.f, 0.f, 0.f, 0.f);
  if (kPhilox) nz_cur = noise_pair(kg, 0u, iter_lo, iter_hi_e, key, sig0, sig1);
  if (kPhilox && kStoch) xi_cur = xi_quad(kg, 0u, iter_lo, iter_hi_e, key);

  // mean sequence and the per-step action-cost coefficients u_prev[t] Sigma^-1 (mppi.py:178-181)
  for (int i = tid; i < 2 * T; i += blockDim.x) {
    const float u = u_prev_e[i];
    uprev_s[i] = u;
    float* uc = coef_s + 4 * (i >> 1) + (i & 1);
    uc[0] = u;
    uc[2] = __fmul_rn(u, (i & 1) ? P.icov1 : P.icov0);
  }
  if (!kRecord && blockIdx.x == 0 && P.replay != nullptr) {  // what a later re-roll of selected samples starts from
    float* rp = P.replay + static_cast<size_t>(env) * (2 * T + 4);
    for (int i = tid; i < 2 * T; i += blockDim.x) rp[i] = u_prev_e[i];
    if (tid == 0) {
      rp[2 * T] = sx;
      rp[2 * T + 1] = sy;
      rp[2 * T + 2] = sth;
    }
  }
  const float4* ucf_s = reinterpret_cast<const float4*>(coef_s);
  __syncthreads();
  if (kPatch) mbar_wait(bar_patch, 0);
  if (!kPhilox && !kWide) {
    if (nz_bulk) mbar_wait(bar_noise + warp, 0);
    else __syncwarp();
  }
  if (!kWide) {  // the latency variant keeps every loop invariant in a register; the wide one must fit 128 registers
    C.pin_all();
    pin(sig0);
    pin(sig1);
  }
  const long long t_loop0 = clock64();

  // ---- T-step rollout, one sample per thread
  float cost = FLT_MAX;
  float* nrow = nz_w + lane * (kWide ? kWideNzStride : 2 * T);
  float* rrow = rec_s + (warp * 32 + lane) * (kWide ? kWideRecStride : 3 * rec_slots);
  // wide variant: flush recorded-state slots [t0, t0 + nt_rec) and noise steps [t0, t0 + nt_nz) of this warp's chunk
  // slabs to HBM (row segments of nt * 12 / nt * 8 bytes: whole sectors except at the segment ends), then rebase the
  // slab pointers so that step t0 + nt lands at the start of the slab.  A full warp copies cooperatively (one row
  // segment per pass, consecutive lanes on consecutive words); the one ragged warp of a launch copies lane by lane.
  const bool nz_vec_ok = kWide && (T & 1) == 0 && (reinterpret_cast<uintptr_t>(noise_out_e) & 15u) == 0u;
  auto flush_chunk = [&](int t0, int nt_rec, int nt_nz) {
    if (!kWide) return;
    float* rec_g = kRecord ? rec_e + static_cast<size_t>(warp_first) * 3 * (T + 1) + 3 * t0 : nullptr;
    float* nz_g = noise_out_e + static_cast<size_t>(warp_first) * 2 * T + 2 * t0;
    if (warp_rows == 32) {
      __syncwarp();
      if (kRecord) {
        const float* src = rec_s + warp * 32 * kWideRecStride;
        if (nt_rec == kChunkSteps) flush_rec_chunk(rec_g, T, src, rec_flush_lane(T, lane));  // (offsets recomputed per
                                                                                             // flush: not kept live in the loop)
        else copy_rows_flat(rec_g, 3 * (T + 1), src, kWideRecStride, 3 * nt_rec, lane);
      }
      if (kPhilox && nt_nz > 0) {
        if (nz_vec_ok && nt_nz == kChunkSteps) {  // 16-byte words: 8 lanes per row, 4 rows per pass
          const int u = lane & 7;
#pragma unroll
          for (int r0 = 0; r0 < 32; r0 += 4) {
            const int r = r0 + (lane >> 3);
            const float4 v = *reinterpret_cast<const float4*>(nz_w + r * kWideNzStride + 4 * u);
            *reinterpret_cast<float4*>(nz_g + static_cast<size_t>(r) * 2 * T + 4 * u) = v;
          }
        } else {
          copy_rows_flat(nz_g, 2 * T, nz_w, kWideNzStride, 2 * nt_nz, lane);
        }
      }
      __syncwarp();
    } else if (valid) {
      if (kRecord) {
        float* dst = rec_g + static_cast<size_t>(lane) * 3 * (T + 1);
        const float* src = rec_s + (warp * 32 + lane) * kWideRecStride;
        for (int i = 0; i < 3 * nt_rec; ++i) dst[i] = src[i];
      }
      if (kPhilox) {
        float* dst = nz_g + static_cast<size_t>(lane) * 2 * T;
        const float* src = nz_w + lane * kWideNzStride;
        for (int i = 0; i < 2 * nt_nz; ++i) dst[i] = src[i];
      }
    }
    rrow -= 3 * nt_nz;
    nrow -= 2 * nt_nz;
  };
  const float* nz_in_row = (kWide && !kPhilox) ? P.noise_in + (eK + k) * 2 * T : nullptr;  // injected noise, wide
  // mid-loop flush of the slab's first half (rec_split): coalesced by the whole warp when it is full, else (the one
  // ragged warp of a launch) every lane writes its own row; afterwards slot t lives at rrow[3 (t - Tc)]
  auto flush_first_half = [&]() {
    if (kRecord) {
      const int Tc = P.rec_split;
      float* rec_g = rec_e + static_cast<size_t>(warp_first) * 3 * (T + 1);
      if (warp_rows == 32) {
        __syncwarp();
        copy_rec_slots(rec_g, rec_s + warp * 32 * 3 * rec_slots, 32, T, rec_slots, 0, Tc, lane);
        __syncwarp();
      } else {
        float* dst = rec_g + static_cast<size_t>(lane) * 3 * (T + 1);
        for (int i = 0; i < 3 * Tc; ++i) dst[i] = rrow[i];
      }
      rrow -= 3 * Tc;
    }
  };
  const int split_pair = (kRecord && P.rec_split > 0) ? (P.rec_split >> 1) : -1;
  if (valid) {
    SampleState s;
    s.x = sx; s.y = sy; s.th = sth; s.stage_sum = 0.0f; s.act0 = 0.0f; s.act1 = 0.0f;
    s.tau = 0.0f;
    s.ms = make_float2(0.0f, 0.0f);
    if (kStoch) s.ms = lookup_slip<kPatch, kPow2, false>(C, s.x, s.y);
    else s.tau = lookup_tau<kPatch, kPow2, false>(C, s.x, s.y);
    const float* xrow = (kStoch && !kPhilox) ? P.xi_in + (eK + k) * (2 * T + 1) : nullptr;
    // Steps are processed in pairs (one Philox call yields both steps' noise).  The pair after the current one is
    // drawn in the same straight-line block as the current pair's steps -- unconditionally, so that there is no
    // branch and ptxas can interleave its ~100 independent instructions into the stall slots of the dependency
    // chain (this warp is alone on its scheduler).  Step 0 takes the general math (arbitrary initial heading);
    // an odd horizon's last step is peeled off the pair loop.
    const int nfull = T >> 1;
    auto fetch_pair = [&](int p) -> float4 {  // noise of steps 2p, 2p+1; advances the Philox pipeline
      float4 nz;
      if (kPhilox) {
        nz = nz_cur;
        nz_cur = noise_pair(kg, static_cast<uint32_t>(p + 1), iter_lo, iter_hi_e, key, sig0, sig1);
      } else if (kWide) {
        const float* g = nz_in_row + 4 * p;
        nz = make_float4(__ldg(g), __ldg(g + 1), __ldg(g + 2), __ldg(g + 3));
      } else {
        const float2 a = *reinterpret_cast<const float2*>(nrow + 4 * p);
        const float2 b = *reinterpret_cast<const float2*>(nrow + 4 * p + 2);
        nz = make_float4(a.x, a.y, b.x, b.y);
      }
      return nz;
    };
    auto fetch_xi = [&](int p) -> float4 {  // lookup normals (transit 2p, stage 2p, transit 2p+1, stage 2p+1)
      float4 xq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kStoch) {
        if (kPhilox) {
          xq = xi_cur;
          xi_cur = xi_quad(kg, static_cast<uint32_t>(p + 1), iter_lo, iter_hi_e, key);
        } else {
          xq = make_float4(__ldg(xrow + 4 * p), __ldg(xrow + 4 * p + 1), __ldg(xrow + 4 * p + 2), __ldg(xrow + 4 * p + 3));
        }
      }
      return xq;
    };
    if (nfull >= 1) {
      const float4 nz = fetch_pair(0);
      const float4 xq = fetch_xi(0);
      if (kPhilox) {
        *reinterpret_cast<float2*>(nrow) = make_float2(nz.x, nz.y);
        *reinterpret_cast<float2*>(nrow + 2) = make_float2(nz.z, nz.w);
      }
      sample_step<kPatch, kPow2, kRecord, false, kStoch>(s, C, 0, nz.x, nz.y, xq.x, xq.y, ucf_s, rrow);
      sample_step<kPatch, kPow2, kRecord, kFastAngles, kStoch>(s, C, 1, nz.z, nz.w, xq.z, xq.w, ucf_s, rrow);
      // The pair loop itself stays branch-free (the codegen of this loop is what the iteration's latency hangs
      // on: a flush test inside it cost 50 cycles per step).  With a split slab it simply runs twice, around the flush.
      auto run_pairs = [&](int p_begin, int p_end) {
        for (int p = p_begin; p < p_end; ++p) {
          const float4 nq = fetch_pair(p);
          const float4 xp = fetch_xi(p);
          if (kPhilox) {
            if (kWide) {
              *reinterpret_cast<float4*>(nrow + 4 * p) = nq;
            } else {
              *reinterpret_cast<float2*>(nrow + 4 * p) = make_float2(nq.x, nq.y);
              *reinterpret_cast<float2*>(nrow + 4 * p + 2) = make_float2(nq.z, nq.w);
            }
          }
          sample_step<kPatch, kPow2, kRecord, kFastAngles, kStoch>(s, C, 2 * p, nq.x, nq.y, xp.x, xp.y, ucf_s, rrow);
          sample_step<kPatch, kPow2, kRecord, kFastAngles, kStoch>(s, C, 2 * p + 1, nq.z, nq.w, xp.z, xp.w, ucf_s, rrow);
        }
      };
      if (kWide) {  // chunks of kChunkPairs step pairs, each flushed as soon as it is complete
        for (int pc = 0; pc < nfull; pc += kChunkPairs) {
          const int pe = min(pc + kChunkPairs, nfull);
          run_pairs(max(pc, 1), pe);
          flush_chunk(2 * pc, 2 * (pe - pc), 2 * (pe - pc));
        }
      } else if (split_pair < 0) {
        run_pairs(1, nfull);
      } else {
        run_pairs(1, split_pair);
        flush_first_half();
        run_pairs(split_pair, nfull);
      }
    }
    if (T & 1) {  // last (or only) step of an odd horizon: first half of pair nfull
      float2 nl;
      float xl_tr = 0.0f, xl_st = 0.0f;
      if (kPhilox) {
        nl = make_float2(nz_cur.x, nz_cur.y);
        *reinterpret_cast<float2*>(nrow + 4 * nfull) = nl;
        xl_tr = xi_cur.x;
        xl_st = xi_cur.y;
      } else {
        nl = kWide ? make_float2(__ldg(nz_in_row + 4 * nfull), __ldg(nz_in_row + 4 * nfull + 1))
                   : *reinterpret_cast<const float2*>(nrow + 4 * nfull);
        if (kStoch) {
          xl_tr = __ldg(xrow + 4 * nfull);
          xl_st = __ldg(xrow + 4 * nfull + 1);
        }
      }
      if (nfull == 0) sample_step<kPatch, kPow2, kRecord, false, kStoch>(s, C, 0, nl.x, nl.y, xl_tr, xl_st, ucf_s, rrow);
      else sample_step<kPatch, kPow2, kRecord, kFastAngles, kStoch>(s, C, T - 1, nl.x, nl.y, xl_tr, xl_st, ucf_s, rrow);
    }
    if (kRecord) { rrow[3 * T + 0] = s.x; rrow[3 * T + 1] = s.y; rrow[3 * T + 2] = s.th; }
    // wide: what is left in the slabs -- an odd horizon's last step and the final (clamped, wrapped) state
    flush_chunk(2 * nfull, T + 1 - 2 * nfull, T - 2 * nfull);
    // terminal cost (mppi.py:184; objectives.py:65): same cell as the last stage cost, its own draw when stochastic
    float tau_term = s.tau;
    if (kStoch) {
      const float xt = kPhilox ? xi_quad(kg, kXiTerminalPair, iter_lo, iter_hi_e, key).x : __ldg(xrow + 2 * T);
      tau_term = slip_to_trav(s.ms, xt);
    }
    const float terminal = goal_and_stuck_cost_at(C, term_gx, term_gy, s.x, s.y, tau_term);
    cost = __fadd_rn(__fadd_rn(s.stage_sum, terminal), __fmul_rn(P.lambda, s.act0 + s.act1));   // mppi.py:186-190
    costs_e[k] = cost;
  }
  const long long t_loop1 = clock64();

  // ---- per-CTA softmax partial: m = max score, s = sum exp(score - m), U[c] = sum exp(score - m) * v[k][c]
  const float score = valid ? score_of(cost, P) : -FLT_MAX;
  float wm = warp_max(score);
  if (lane == 0) red_s[warp] = wm;
  __syncthreads();
  float m_cta = red_s[0];
  for (int w = 1; w < nwarps; ++w) m_cta = fmaxf(m_cta, red_s[w]);
  const float e = valid ? __expf(score - m_cta) : 0.0f;
  e_s[tid] = e;
  float ws = warp_sum(e);
  if (lane == 0) red_s[8 + warp] = ws;
  __syncwarp();
  // each warp: columns over lanes, its own 32 samples; the clamped controls v = clamp(u_prev + noise) are rebuilt from
  // the noise (the same two ops as in sample_step), e_s holds the un-normalised weights.  Row r always goes to
  // accumulator r mod 4, in ascending order, in both variants.
  if (kWide) {
    // the noise left shared memory chunk by chunk: re-read this warp's rows from L2 / HBM, lanes on consecutive columns
    // (coalesced), 16 rows x 4 column groups = 64 loads in flight per lane.  A batch of rows whose un-normalised weights
    // all underflowed to exactly 0 is skipped (its exact zeros would not change a bit of the sums; a NaN weight counts
    // as non-zero and propagates).
    const unsigned nzmask = __ballot_sync(0xffffffffu, e != 0.0f);
    const float* ew = e_s + warp * 32;
    const int ncol2 = 2 * T;
    const float* nz_rows = (kPhilox ? noise_out_e : P.noise_in + eK * 2 * T) + static_cast<size_t>(warp_first) * 2 * T;
    const float lo = (lane & 1) ? C.u_min1 : C.u_min0, hi = (lane & 1) ? C.u_max1 : C.u_max0;  // column parity = lane parity
    for (int c0 = 0; c0 < ncol2; c0 += 128) {
      float acc[4][4];
      float um[4];
      int cc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cc[i] = c0 + 32 * i + lane;
        um[i] = cc[i] < ncol2 ? uprev_s[cc[i]] : 0.0f;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q][i] = 0.0f;
      }
      const float* colp[4];  // this lane's four columns; a column group past the last column reads column 0 and is
      bool col_ok[4];        // discarded at the end
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        col_ok[i] = cc[i] < ncol2;
        colp[i] = nz_rows + (col_ok[i] ? cc[i] : 0);
      }
#pragma unroll 1
      for (int r0 = 0; r0 < 32; r0 += 16) {
        if (((nzmask >> r0) & 0xFFFFu) == 0u) continue;
        float v[16][4];
        if (warp_rows == 32) {  // full warp: no row guards, row offsets are warp-uniform
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int roff = (r0 + j) * ncol2;
#pragma unroll
            for (int i = 0; i < 4; ++i) v[j][i] = __ldcg(colp[i] + roff);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              v[j][i] = (r0 + j < warp_rows && col_ok[i]) ? __ldcg(colp[i] + (r0 + j) * ncol2) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float ej = ew[r0 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[j & 3][i] = fmaf(ej, clampf(__fadd_rn(um[i], v[j][i]), lo, hi), acc[j & 3][i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (cc[i] < ncol2) warpu_s[warp * ncol2 + cc[i]] = (acc[0][i] + acc[1][i]) + (acc[2][i] + acc[3][i]);
    }
  } else {
    const float4* e4 = reinterpret_cast<const float4*>(e_s + warp * 32);
    for (int c = lane; c < 2 * T; c += 32) {
      const float* col = nz_w + c;
      const float um = uprev_s[c];
      const float lo = (c & 1) ? C.u_min1 : C.u_min0, hi = (c & 1) ? C.u_max1 : C.u_max0;
      float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
#pragma unroll
      for (int r4 = 0; r4 < 8; ++r4) {  // rows past warp_rows carry e = 0 and zero noise
        const float4 ev = e4[r4];
        acc0 = fmaf(ev.x, clampf(__fadd_rn(um, col[(4 * r4 + 0) * 2 * T]), lo, hi), acc0);
        acc1 = fmaf(ev.y, clampf(__fadd_rn(um, col[(4 * r4 + 1) * 2 * T]), lo, hi), acc1);
        acc2 = fmaf(ev.z, clampf(__fadd_rn(um, col[(4 * r4 + 2) * 2 * T]), lo, hi), acc2);
        acc3 = fmaf(ev.w, clampf(__fadd_rn(um, col[(4 * r4 + 3) * 2 * T]), lo, hi), acc3);
      }
      warpu_s[warp * 2 * T + c] = (acc0 + acc1) + (acc2 + acc3);
    }
  }
  __syncthreads();
  float s_cta = 0.0f;
  for (int w = 0; w < nwarps; ++w) s_cta += red_s[8 + w];
  for (int c = tid; c < 2 * T; c += blockDim.x) {
    float acc = 0.0f;
    for (int w = 0; w < nwarps; ++w) acc += warpu_s[w * 2 * T + c];
    part_u_e[static_cast<size_t>(blockIdx.x) * 2 * T + c] = acc;
  }
  if (tid == 0) {
    part_ms_e[2 * blockIdx.x + 0] = m_cta;
    part_ms_e[2 * blockIdx.x + 1] = s_cta;
  }
  const long long t_part = clock64();
  // Two epilogue schedules.  coop (the whole grid is co-resident; cooperative launch): the slab stores (17 MB over
  // the grid) are held back until the last CTA has merged the partials -- issued earlier, their burst through
  // L2 was measured to stretch the merge's loads from ~0.8k to ~3k cycles each -- and every CTA normalises its own
  // weights from registers once (M, S) are published.  Otherwise (more CTAs than SMs): stores go out at once and
  // the last CTA rescales all weights.
  const bool coop = P.coop != 0;
  if (!coop) {
    if (valid) weights_e[k] = e;  // exp(score - m_cta); rescaled by the last CTA
    if (!kWide)
      store_slabs<kRecord, kPhilox>(P, rec_e, noise_out_e, rec_s, nz_w, warp, lane, warp_first, warp_rows, nz_bytes);
  }

  // ---- grid-wide merge (per environment) by the last CTA to finish (atomic ticket).  The CTA barrier orders every
  // thread's writes before thread 0's acq_rel atomic (cumulativity): the release half publishes this CTA's partial,
  // the acquire half (in the CTA that draws the last ticket) makes every other CTA's partial visible to the loads below.
  // Grids of more than kTwoLevelMin CTAs (several waves: the wide variant at K >= 32768) merge in two levels: the
  // last CTA of every group of kMergeGroup consecutive CTAs merges its group's partials into one group partial, the
  // last group to finish merges those -- each merge reads a few KB instead of the whole array (512 CTAs: 209 KB into
  // one SM, measured 12 us), and the group merges overlap the rollouts still running.
  const int nblk_all = gridDim.x;
  const int ncol = 2 * T;
  const bool two_level = !coop && nblk_all > kTwoLevelMin;  // (co-resident grids: measured SLOWER in two levels --
                                                            // 256 CTAs: 53.1 vs 47.6 us at config 4, 42.3 vs 36.7 us
                                                            // deterministic -- the waiting CTAs make the extra ticket
                                                            // round trips expensive)
  const int ngroups = (nblk_all + kMergeGroup - 1) / kMergeGroup;
  const int my_group = blockIdx.x / kMergeGroup;
  const int group_cnt = min(kMergeGroup, nblk_all - my_group * kMergeGroup);
  unsigned int* const ticket_grp_e = P.ticket_grp + static_cast<size_t>(env) * P.max_groups;
  float* const part2_ms_e = P.part2_ms + static_cast<size_t>(env) * P.max_groups * 2;
  float* const part2_u_e = P.part2_u + static_cast<size_t>(env) * P.max_groups * ncol;
  __syncthreads();
  if (tid == 0) {
    if (two_level) {
      const unsigned int prev = atom_add_acq_rel_gpu(ticket_grp_e + my_group, 1u);
      *last_flag = (prev == static_cast<unsigned int>(group_cnt) - 1u) ? 2 : 0;
      if (*last_flag) ticket_grp_e[my_group] = 0u;  // every member has arrived: re-arm for the next launch
    } else {
      const unsigned int prev = atom_add_acq_rel_gpu(ticket_e, 1u);
      *last_flag = (prev == gridDim.x - 1) ? 1 : 0;
    }
  }
  __syncthreads();
  int role = *last_flag;  // 0 = done after the partial, 1 = merges the grid (or the groups), 2 = merges its group
  float M = 0.0f, S = 1.0f;
  const bool stamp = P.dbg_ts != nullptr && env == 0;
  float* a_s = reinterpret_cast<float*>(smem + L.off_merge);  // per-CTA rescale factors exp(m_g - M)
  float* grp_s = a_s + kMergeACap;                             // [ngrp][ncol] partial column sums
  // LSE merge of `nblk` partials (ms_src [nblk][2], u_src [nblk][2T]) by this CTA: leaves M, S and the un-normalised
  // column sums U in uprev_s.  Called by every thread of the CTA.
  auto merge_partials = [&](const float* part_ms_e, const float* part_u_e, const int nblk) {
    // Fast path (2T a multiple of 4, one float4 column unit per thread): everything the merge needs from other SMs --
    // all (m_g, s_g) and this thread's share of the U_g rows -- is requested up front and consumed from registers, so
    // the merge costs one L2 round trip.  EVERY WARP computes M, S and the rescale factors a_g = exp(m_g - M) for itself
    // (same lanes, same order: bit-identical in every warp) into its own copy in shared memory, so the only CTA
    // barrier of the merge is the one between the scaled row sums and the column sums (round 1: four barriers,
    // 3.5k cycles after the loads; now ~1.5k).  Fixed assignment and fixed-order sums keep the result bit-reproducible.
    constexpr int kMsCache = 10, kMergeBatch = kWide ? 8 : 32;  // (the wide variant lives within 128 registers)
    const int nunit = ncol >> 2;
    const bool fast_merge = (ncol & 3) == 0 && nunit <= static_cast<int>(blockDim.x) && nblk * nwarps <= kMergeACap &&
                            ncol <= kMergeGrpCap && nblk <= kMsCache * 32;
    int ngrp = 1;
    bool s_known = false;
    if (fast_merge) {
      ngrp = min(min(static_cast<int>(blockDim.x) / nunit, kMergeGrpCap / ncol), nblk);
      float2 ms[kMsCache];
#pragma unroll
      for (int j = 0; j < kMsCache; ++j) {
        const int g = lane + 32 * j;
        ms[j] = make_float2(-FLT_MAX, 0.0f);
        if (g < nblk) ms[j] = __ldcg(reinterpret_cast<const float2*>(part_ms_e) + g);
      }
      const int unit = tid % nunit, grp = tid / nunit;
      const int cnt = (grp < ngrp) ? (nblk - grp + ngrp - 1) / ngrp : 0;  // CTAs grp, grp + ngrp, ... owned by this thread
      const float4* src = reinterpret_cast<const float4*>(part_u_e + static_cast<size_t>(grp) * ncol) + unit;
      const size_t stride4 = static_cast<size_t>(ngrp) * nunit;
      float4 pv[kMergeBatch];
#pragma unroll
      for (int j = 0; j < kMergeBatch; ++j) {
        pv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < cnt) pv[j] = __ldcg(src + j * stride4);
      }
      float lm = -FLT_MAX;
#pragma unroll
      for (int j = 0; j < kMsCache; ++j) lm = fmaxf(lm, ms[j].x);
      if (stamp) BNV_STAMP(12);
      M = warp_max(lm);
      float* a_w = a_s + warp * nblk;  // this warp's copy of the rescale factors
      float lsum = 0.0f;
#pragma unroll
      for (int j = 0; j < kMsCache; ++j) {
        const int g = lane + 32 * j;
        if (g < nblk) {
          const float a = __expf(ms[j].x - M);
          a_w[g] = a;
          lsum = fmaf(a, ms[j].y, lsum);
        }
      }
      S = warp_sum(lsum);
      s_known = true;
      __syncwarp();
      if (stamp) BNV_STAMP(14);
      if (cnt > 0) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* ap = a_w + grp;
#pragma unroll
        for (int j = 0; j < kMergeBatch; ++j) {
          if (j < cnt) {
            const float a = ap[j * ngrp];
            acc.x = fmaf(a, pv[j].x, acc.x);
            acc.y = fmaf(a, pv[j].y, acc.y);
            acc.z = fmaf(a, pv[j].z, acc.z);
            acc.w = fmaf(a, pv[j].w, acc.w);
          }
        }
        for (int j0 = kMergeBatch; j0 < cnt; j0 += 8) {  // more CTAs than the first batch covers: eight loads at a time
                                                         // (whole batches of 32 again measured slower: 15k vs 7k cycles
                                                         // for the 20 extra rows per thread of a 256-CTA grid)
          float4 q[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            q[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j0 + jj < cnt) q[jj] = __ldcg(src + (j0 + jj) * stride4);
          }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            if (j0 + jj < cnt) {
              const float a = ap[(j0 + jj) * ngrp];
              acc.x = fmaf(a, q[jj].x, acc.x);
              acc.y = fmaf(a, q[jj].y, acc.y);
              acc.z = fmaf(a, q[jj].z, acc.z);
              acc.w = fmaf(a, q[jj].w, acc.w);
            }
          }
        }
        *reinterpret_cast<float4*>(grp_s + grp * ncol + unit * 4) = acc;
      }
    } else {
      // general path (odd horizons, very long horizons, very many CTAs): plain loops, a_g recomputed on the fly
      float lm = -FLT_MAX;
      for (int g = tid; g < nblk; g += blockDim.x) lm = fmaxf(lm, __ldcg(part_ms_e + 2 * g));
      lm = warp_max(lm);
      if (lane == 0) red_s[16 + warp] = lm;
      __syncthreads();
      M = red_s[16];
      for (int w = 1; w < nwarps; ++w) M = fmaxf(M, red_s[16 + w]);
      float lsum = 0.0f;
      for (int g = tid; g < nblk; g += blockDim.x)
        lsum = fmaf(__expf(__ldcg(part_ms_e + 2 * g) - M), __ldcg(part_ms_e + 2 * g + 1), lsum);
      lsum = warp_sum(lsum);
      if (lane == 0) red_s[24 + warp] = lsum;
      for (int c = tid; c < ncol; c += blockDim.x) {
        float acc = 0.0f;
        for (int g = 0; g < nblk; ++g)
          acc = fmaf(__expf(__ldcg(part_ms_e + 2 * g) - M), __ldcg(part_u_e + static_cast<size_t>(g) * ncol + c), acc);
        grp_s[c] = acc;
      }
    }
    if (stamp) BNV_STAMP(15);
    __syncthreads();
    if (!s_known) {
      S = 0.0f;
      for (int w = 0; w < nwarps; ++w) S += red_s[24 + w];
    }
    for (int c = tid; c < ncol; c += blockDim.x) {
      float acc = 0.0f;
      for (int gq = 0; gq < ngrp; ++gq) acc += grp_s[gq * ncol + c];
      uprev_s[c] = acc;  // U of these partials, un-normalised
    }
    __syncthreads();
  };
  if (role != 0 && stamp && tid == 0) {
    P.dbg_ts[0] = t_start; P.dbg_ts[1] = t_loop0; P.dbg_ts[2] = t_loop1; P.dbg_ts[3] = t_part; P.dbg_ts[4] = clock64();
  }
  if (role == 2) {
    merge_partials(part_ms_e + 2 * my_group * kMergeGroup, part_u_e + static_cast<size_t>(my_group) * kMergeGroup * ncol,
                   group_cnt);
    if (tid == 0) {
      part2_ms_e[2 * my_group + 0] = M;
      part2_ms_e[2 * my_group + 1] = S;
    }
    for (int c = tid; c < ncol; c += blockDim.x) part2_u_e[static_cast<size_t>(my_group) * ncol + c] = uprev_s[c];
    __syncthreads();
    if (tid == 0) {
      const unsigned int prev = atom_add_acq_rel_gpu(ticket_e, 1u);
      *last_flag = (prev == static_cast<unsigned int>(ngroups) - 1u) ? 1 : 0;
    }
    __syncthreads();
    role = *last_flag;
    if (role == 1) merge_partials(part2_ms_e, part2_u_e, ngroups);
  } else if (role == 1) {
    merge_partials(part_ms_e, part_u_e, nblk_all);
  }
  const bool is_last = role == 1;
  if (is_last) {
    // ---- sharded softmax, fused exchange over NVLink peer memory (exchange_column): one thread per column
    const bool fused = !kBatch && P.world > 1 && P.peer_mbox != nullptr;
    if (fused) {
      for (int c = tid; c < ncol; c += blockDim.x) {
        const ColumnMerge cm = exchange_column(P, c, ncol, uprev_s[c], M, S);
        uprev_s[c] = cm.U;
        if (c == 0) {
          red_s[34] = cm.M;
          red_s[35] = cm.S;
        }
      }
      __syncthreads();
      M = red_s[34];
      S = red_s[35];
    }
    const bool complete = P.world == 1 || fused;  // (M, S, U) now cover every sample of the solver
    // publish (M, S) and release the waiting CTAs as early as possible -- from the LAST thread: the release is a
    // membar that stalls its warp for several hundred cycles, and warp 0 (thread 0 runs the serial optimal rollout
    // next) must not wait for it.  So with more than one warp the last warp only publishes, the other warps compute
    // u* and meet at a named barrier of their own.
    const bool split_pub = coop && blockDim.x > 32;
    const int nwork = split_pub ? static_cast<int>(blockDim.x) - 32 : static_cast<int>(blockDim.x);
    if (coop && tid == static_cast<int>(blockDim.x) - 1) {
      stats_e[0] = M;
      stats_e[1] = S;
      st_release_gpu(ticket_e + 1, epoch);
    }
    if (complete) {
      float* u_out_e = P.u_out + static_cast<size_t>(env) * ncol;
      if (tid < nwork) {
        for (int c = tid; c < ncol; c += nwork) {
          const float u = __fdiv_rn(uprev_s[c], S);  // u* = U / S (mppi.py:196-199)
          uprev_s[c] = u;
          warpu_s[c] = clampf(u, (c & 1) ? C.u_min1 : C.u_min0, (c & 1) ? C.u_max1 : C.u_max0);  // for the optimal rollout
          u_out_e[c] = u;
          if (P.keep_mean) u_prev_e[c] = u;  // next call's mean sequence, unshifted (mppi.py:217)
        }
      }
    } else {  // unfused sharding: hand the shard partial to the host-side exchange + finalize_kernel
      if (tid == 0) {
        P.shard_partial[0] = M;
        P.shard_partial[1] = S;
      }
      for (int c = tid; c < ncol; c += blockDim.x) P.shard_partial[2 + c] = uprev_s[c];
    }
    if (split_pub) {
      if (tid < nwork) asm volatile("bar.sync 1, %0;" ::"r"(nwork) : "memory");
    } else {
      __syncthreads();
    }
    if (complete && tid == nwork - 1) signal_action(P);  // (thread 0 goes straight on to the optimal rollout)
    if (stamp) BNV_STAMP(5);
    if (tid == 0) *ticket_e = 0u;  // re-arm for the next launch (all arrivals of this one have happened)
    if (stamp) BNV_STAMP(6);
    if (!coop) {
      // (M, S) for normalize_weights_kernel, launched right behind this kernel: weights[k] = exp(score_k - m_cta) *
      // exp(m_cta - M) / S (softmax, mppi.py:193; 1/S deferred to finalize_kernel when unfused-sharded).  Rescaling
      // all K weights here, by one CTA, was the longest serial piece of a large launch (1 MB through one SM).
      if (tid == 32 % blockDim.x) {
        stats_e[0] = M;
        stats_e[1] = complete ? S : 1.0f;
      }
      if (complete && warp == 0) {
        const float* xi_row = (kStoch && !kPhilox) ? P.xi_opt_in + static_cast<size_t>(env) * T : nullptr;
        if (kStoch && kPhilox && T <= kMergeGrpCap && !(P.dbg_flags & 2u)) {  // (grp_s is free once the merge is done)
          predraw_optimal_xi(grp_s, T, iter_lo, iter_hi_e, key, lane);
          xi_row = grp_s;
        }
        if (lane == 0) {
          if (kWide) C.pin_all();  // the serial rollout is a pure latency chain: its invariants belong in registers
          const XiSource xs{xi_row, kOptimalSample, iter_lo, iter_hi_e, key};
          optimal_rollout<kPatch, kPow2, kFastAngles, kStoch>(T, C, warpu_s, sx, sy, sth,
                                                              P.opt_rec + static_cast<size_t>(env) * 3 * (T + 1), xs);
          signal_done(P);
        }
      }
      if (stamp) BNV_STAMP(7);
    }
  } else if (coop) {
    // wait for the last CTA's merge (all CTAs are co-resident: cooperative launch), then pick up (M, S)
    if (tid == 0) {
      while (ld_acquire_gpu(ticket_e + 1) != epoch) __nanosleep(32);
    }
    __syncthreads();
    M = __ldcg(stats_e);
    S = __ldcg(stats_e + 1);
  }
  if (coop) {
    // softmax weight of this thread's sample straight from registers (mppi.py:193; 1/S deferred when unfused-sharded)
    const bool complete = P.world == 1 || (!kBatch && P.peer_mbox != nullptr);
    if (valid) weights_e[k] = e * __expf(m_cta - M) * (complete ? __fdiv_rn(1.0f, S) : 1.0f);
    // (measured, not adopted: the last CTA holding its OWN slab stores back until the serial optimal rollout is done --
    // the rollout runs unobstructed, 13.7k -> 7.1k cycles at two CTAs per SM, 7.1k -> 6.7k at one, but the CTA's
    // slab then drains alone at the SM's ~30 B/clk store rate behind it: 23.2 -> 25.2 us at configuration 1)
    if (!kWide)
      store_slabs<kRecord, kPhilox>(P, rec_e, noise_out_e, rec_s, nz_w, warp, lane, warp_first, warp_rows, nz_bytes);
    if (is_last && complete && warp == 0) {
      const float* xi_row = (kStoch && !kPhilox) ? P.xi_opt_in + static_cast<size_t>(env) * T : nullptr;
      if (kStoch && kPhilox && T <= kMergeGrpCap && !(P.dbg_flags & 2u)) {
        predraw_optimal_xi(grp_s, T, iter_lo, iter_hi_e, key, lane);
        xi_row = grp_s;
      }
      if (lane == 0) {
        const XiSource xs{xi_row, kOptimalSample, iter_lo, iter_hi_e, key};
        optimal_rollout<kPatch, kPow2, kFastAngles, kStoch>(T, C, warpu_s, sx, sy, sth,
                                                            P.opt_rec + static_cast<size_t>(env) * 3 * (T + 1), xs);
        signal_done(P);
      }
    }
    if (is_last && P.dbg_ts != nullptr && env == 0) BNV_STAMP(7);
  }
  // shared memory must stay allocated until the bulk stores have read it
  if (!kWide && (kRecord || kPhilox)) bulk_wait_read_all();
  if (is_last && P.dbg_ts != nullptr && env == 0) BNV_STAMP_ANY(10, 32);
}

#ifndef BNV_ROLLOUT_ONLY
// --------------------------------------------------------------------------------------------- finalize (multi-GPU)
// gathered: [world][2+2T] shard partials, identical on every rank -> every rank computes the same u*.
// One CTA: merges, rescales this shard's weights (they hold exp(score - M_shard)), writes the outputs and runs
// the optimal rollout.
template <bool kPatch, bool kPow2, bool kFastAngles>
__global__ void __launch_bounds__(kFinalizeThreads) finalize_kernel(const __grid_constant__ EngineParams P,
                                                                    const float* __restrict__ gathered, int rank) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int T = P.T, ncol = 2 * T, plen = 2 + 2 * T;
  uint64_t* bar_patch = reinterpret_cast<uint64_t*>(smem);
  float* patch_s = reinterpret_cast<float*>(smem + 128);
  float* u_s = patch_s + (kPatch ? ((P.patch_w * P.patch_h + 31) / 32) * 32 : 0);
  if (tid == 0) {
    mbar_init(bar_patch, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const float sx = P.state_inline ? P.state_val[0] : __ldg(P.state);
  const float sy = P.state_inline ? P.state_val[1] : __ldg(P.state + 1);
  const float sth = P.state_inline ? P.state_val[2] : __ldg(P.state + 2);
  const WindowGeom wg = window_for_state(P, sx, sy);
  if (kPatch) {
    if (tid < 32) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_patch, static_cast<uint32_t>(P.patch_w * P.patch_h * 4));
        tma_load_2d(patch_s, &P.tau_map, wg.ox, wg.oy, bar_patch);
      }
    }
  }
  StepConsts C = make_step_consts(P, wg, patch_s, P.tau, P.goal_x, P.goal_y, 1);
  float M = -FLT_MAX;
  for (int g = 0; g < P.world; ++g) M = fmaxf(M, gathered[g * plen]);
  float S = 0.0f;
  for (int g = 0; g < P.world; ++g) S = fmaf(__expf(gathered[g * plen] - M), gathered[g * plen + 1], S);
  for (int c = tid; c < ncol; c += blockDim.x) {
    float acc = 0.0f;
    for (int g = 0; g < P.world; ++g) acc = fmaf(__expf(gathered[g * plen] - M), gathered[g * plen + 2 + c], acc);
    u_s[c] = acc;
  }
  const float w_scale = __fdiv_rn(__expf(gathered[rank * plen] - M), S);
  __syncthreads();
  if (kPatch) mbar_wait(bar_patch, 0);
  C.pin_all();
  finish_iteration<kPatch, kPow2, kFastAngles>(P, C, S, w_scale, u_s, sx, sy, sth);
}

// --------------------------------------------------------------------------------------------- re-roll (lean solvers)
// get_top_samples of a solver that keeps no recorded states (record_states=False): the n selected samples are rolled
// out AGAIN from what the iteration started from (state and mean sequence saved by the rollout kernel, the sample's
// noise row from the noise array) with the very step function of the rollout kernel -- same operations in the same
// order on the same inputs, so the rows are bit-identical to what a recording solver would have stored (the lookup
// goes to the global map instead of the staged window: same cells).  One thread = one selected sample.
template <bool kPow2, bool kFastAngles>
__global__ void __launch_bounds__(128) reroll_kernel(const __grid_constant__ EngineParams P, const float* __restrict__ noise,
                                                     const float* __restrict__ replay, const int* __restrict__ idx, int n,
                                                     float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* coef_s = reinterpret_cast<float*>(smem);
  const int T = P.T;
  for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) {
    const float u = replay[i];
    float* uc = coef_s + 4 * (i >> 1) + (i & 1);
    uc[0] = u;
    uc[2] = __fmul_rn(u, (i & 1) ? P.icov1 : P.icov0);
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4* ucf_s = reinterpret_cast<const float4*>(coef_s);
  StepConsts C;
  C.x_min = P.geom.x_min; C.y_min = P.geom.y_min; C.x_max = P.geom.x_max; C.y_max = P.geom.y_max;
  C.res = P.geom.res; C.inv_res = P.geom.inv_res; C.dt = P.bounds.dt;
  C.gx = P.goals != nullptr ? __ldg(P.goals) : P.goal_x;
  C.gy = P.goals != nullptr ? __ldg(P.goals + 1) : P.goal_y;
  C.thr = P.thr;
  C.u_min0 = P.bounds.u_min0; C.u_min1 = P.bounds.u_min1; C.u_max0 = P.bounds.u_max0; C.u_max1 = P.bounds.u_max1;
  C.lo_x = 0; C.lo_y = 0; C.hi_x = P.G - 1; C.hi_y = P.G - 1;
  C.pitch = P.pitch;
  C.win_addr = 0u;
  C.map = P.tau;
  C.finish(4u);
  const int k = idx[i];
  const float* nrow = noise + static_cast<size_t>(k) * 2 * T;
  float* rrow = out + static_cast<size_t>(i) * 3 * (T + 1);
  SampleState s;
  s.x = replay[2 * T]; s.y = replay[2 * T + 1]; s.th = replay[2 * T + 2];
  s.stage_sum = 0.0f; s.act0 = 0.0f; s.act1 = 0.0f;
  s.ms = make_float2(0.0f, 0.0f);
  s.tau = lookup_tau<false, kPow2, false>(C, s.x, s.y);
  sample_step<false, kPow2, true, false, false>(s, C, 0, __ldg(nrow), __ldg(nrow + 1), 0.0f, 0.0f, ucf_s, rrow);
  for (int t = 1; t < T; ++t)
    sample_step<false, kPow2, true, kFastAngles, false>(s, C, t, __ldg(nrow + 2 * t), __ldg(nrow + 2 * t + 1), 0.0f, 0.0f,
                                                        ucf_s, rrow);
  rrow[3 * T + 0] = s.x;
  rrow[3 * T + 1] = s.y;
  rrow[3 * T + 2] = s.th;
}

// --------------------------------------------------------------------------------------------- top-n
// MPPI.get_top_samples (mppi.py:221-240).  One CTA: 4-pass byte-wise radix select of the n-th largest
// weight (weights are >= 0, so their bit patterns order like unsigned integers), compaction of the n
// winners, bitonic sort (descending weight, ascending index among ties) and output.  `pairs` is a scratch
// of n_pad (power of two >= n) 64-bit words in shared memory (n_pad <= kTopnSmemPairs) or global memory.
constexpr int kTopnThreads = 1024;
constexpr int kTopnSmemPairs = 16384;

// `rec` != nullptr (small n): the CTA also gathers the selected rows (rec [K][row_len] -> states_out [n][row_len]),
// which saves the separate gather launch; large n leave the gather to gather_rows_kernel (one CTA per row).
constexpr int kTopnFusedGatherMax = 128;
__global__ void __launch_bounds__(kTopnThreads) topn_select_kernel(const float* __restrict__ weights, int K, int n,
                                                                   int n_pad, unsigned long long* pairs_global,
                                                                   float* __restrict__ out_w, int* __restrict__ out_idx,
                                                                   const float* __restrict__ rec, int row_len,
                                                                   float* __restrict__ states_out) {
  extern __shared__ __align__(128) unsigned char smem[];
  // blockIdx.x = environment (batch mode): each CTA selects within its own [K] weights
  weights += static_cast<size_t>(blockIdx.x) * K;
  out_w += static_cast<size_t>(blockIdx.x) * n;
  out_idx += static_cast<size_t>(blockIdx.x) * n;
  if (pairs_global) pairs_global += static_cast<size_t>(blockIdx.x) * n_pad;
  __shared__ unsigned int hist[256];
  __shared__ unsigned int sel_prefix, sel_remaining, n_gt, n_eq;
  unsigned long long* pairs = pairs_global ? pairs_global : reinterpret_cast<unsigned long long*>(smem);
  const int tid = threadIdx.x;
  if (tid == 0) {
    sel_prefix = 0u;
    sel_remaining = static_cast<unsigned int>(n);
    n_gt = 0u;
    n_eq = 0u;
  }
  // The keys are read in tiles of kTile per thread with all of a tile's loads in flight (one L2 round trip per tile
  // instead of one per element: the element-by-element loop cost 16 dependent round trips per pass, 5 passes); when
  // the whole array fits one tile (K <= 16384) it is read ONCE and every pass runs from registers.
  constexpr int kTile = 16;
  const bool cached = K <= kTile * static_cast<int>(blockDim.x);
  unsigned int ck[kTile];
#pragma unroll
  for (int j = 0; j < kTile; ++j) {
    const int i = j * static_cast<int>(blockDim.x) + tid;
    ck[j] = (cached && i < K) ? __float_as_uint(weights[i]) : 0u;
  }
  const int lane = tid & 31;
  unsigned int mask = 0u;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const unsigned int prefix = sel_prefix;
    // (plain shared-memory atomics: grouping a warp's lanes by bucket first -- match.any, or a leader loop over
    // ballots -- was measured slower, 65-69 vs 44 us per call at K = 16384: the weights of a real iteration spread over
    // dozens of exponent buckets, so there is little same-address serialisation to remove)
    for (int base = 0; base < K; base += kTile * static_cast<int>(blockDim.x)) {
      unsigned int key[kTile];
#pragma unroll
      for (int j = 0; j < kTile; ++j) {
        const int i = base + j * static_cast<int>(blockDim.x) + tid;
        key[j] = cached ? ck[j] : (i < K ? __float_as_uint(weights[i]) : 0u);
      }
#pragma unroll
      for (int j = 0; j < kTile; ++j) {
        const int i = base + j * static_cast<int>(blockDim.x) + tid;
        if (i < K && (key[j] & mask) == prefix) atomicAdd(&hist[(key[j] >> shift) & 0xFFu], 1u);
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int remaining = sel_remaining, above = 0u;
      int d = 255;
      for (; d > 0; --d) {
        if (above + hist[d] >= remaining) break;
        above += hist[d];
      }
      sel_remaining = remaining - above;  // how many to take from bucket d (and, finally, among exact ties)
      sel_prefix = prefix | (static_cast<unsigned int>(d) << shift);
    }
    mask |= 0xFFu << shift;
    __syncthreads();
  }
  const unsigned int thr_key = sel_prefix, take_eq = sel_remaining;
  const unsigned int first_eq = static_cast<unsigned int>(n) - take_eq;
  for (int base = 0; base < K; base += kTile * static_cast<int>(blockDim.x)) {  // compaction: one atomic per warp and class
    unsigned int key[kTile];
#pragma unroll
    for (int j = 0; j < kTile; ++j) {
      const int i = base + j * static_cast<int>(blockDim.x) + tid;
      key[j] = cached ? ck[j] : (i < K ? __float_as_uint(weights[i]) : 0u);
    }
#pragma unroll
    for (int j = 0; j < kTile; ++j) {
      const int i = base + j * static_cast<int>(blockDim.x) + tid;
      const unsigned long long packed = (static_cast<unsigned long long>(key[j]) << 32) | static_cast<unsigned int>(~i);
      const bool gt = i < K && key[j] > thr_key, eq = i < K && key[j] == thr_key;
      const unsigned int m_gt = __ballot_sync(0xffffffffu, gt), m_eq = __ballot_sync(0xffffffffu, eq);
      const unsigned int lane_lt = (1u << lane) - 1u;
      unsigned int base_gt = 0u, base_eq = 0u;
      if (lane == 0) {
        if (m_gt) base_gt = atomicAdd(&n_gt, static_cast<unsigned int>(__popc(m_gt)));
        if (m_eq) base_eq = atomicAdd(&n_eq, static_cast<unsigned int>(__popc(m_eq)));
      }
      base_gt = __shfl_sync(0xffffffffu, base_gt, 0);
      base_eq = __shfl_sync(0xffffffffu, base_eq, 0);
      if (gt) pairs[base_gt + __popc(m_gt & lane_lt)] = packed;
      if (eq) {
        const unsigned int slot = base_eq + __popc(m_eq & lane_lt);
        if (slot < take_eq) pairs[first_eq + slot] = packed;
      }
    }
  }
  for (int i = n + tid; i < n_pad; i += blockDim.x) pairs[i] = 0ull;  // pads sort last
  __syncthreads();
  for (int size = 2; size <= n_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (n_pad >> 1); i += blockDim.x) {
        int lo = 2 * i - (i & (stride - 1));
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        unsigned long long a = pairs[lo], b = pairs[hi];
        if ((a < b) == desc) {
          pairs[lo] = b;
          pairs[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n; i += blockDim.x) {
    unsigned long long p = pairs[i];
    out_w[i] = __uint_as_float(static_cast<unsigned int>(p >> 32));
    out_idx[i] = static_cast<int>(~static_cast<unsigned int>(p & 0xFFFFFFFFull));
  }
  if (rec != nullptr) {  // fused gather: one warp per row, lanes on consecutive words
    rec += static_cast<size_t>(blockIdx.x) * K * row_len;
    states_out += static_cast<size_t>(blockIdx.x) * n * row_len;
    const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    for (int i = warp; i < n; i += nw) {
      const unsigned int k = ~static_cast<unsigned int>(pairs[i] & 0xFFFFFFFFull);
      const float* src = rec + static_cast<size_t>(k) * row_len;
      float* dst = states_out + static_cast<size_t>(i) * row_len;
      for (int j = lane; j < row_len; j += 32) dst[j] = src[j];
    }
  }
}

// out[e][i][:] = rec[e][idx[e][i]][:], row = 3 (T+1) floats; blockIdx.y = environment e with K rows each.
__global__ void gather_rows_kernel(const float* __restrict__ rec, const int* __restrict__ idx, int row_len, int K,
                                   float* __restrict__ out) {
  const size_t e = blockIdx.y, n = gridDim.x;
  const float* src = rec + (e * K + idx[e * n + blockIdx.x]) * row_len;
  float* dst = out + (e * n + blockIdx.x) * row_len;
  for (int i = threadIdx.x; i < row_len; i += blockDim.x) dst[i] = src[i];
}

// out[i][:] = table[idx[i]][:row_len], table rows `row_stride` floats apart (bnv_mppi_merge_top: candidate rows are
// {weight, states...}).
__global__ void gather_strided_rows_kernel(const float* __restrict__ table, const int* __restrict__ idx, int row_len,
                                           int row_stride, float* __restrict__ out) {
  const float* src = table + static_cast<size_t>(idx[blockIdx.x]) * row_stride;
  float* dst = out + static_cast<size_t>(blockIdx.x) * row_len;
  for (int i = threadIdx.x; i < row_len; i += blockDim.x) dst[i] = src[i];
}

// --------------------------------------------------------------------------------------------- debug
// L2 flush for measurements: writes `n16` 16-byte words.  Launched with the same dynamic shared-memory size as the
// rollout kernel, so that it can be used to test whether the shared-memory carve-out switch between a flush kernel
// and the rollout kernel is part of the event-timed launch floor (scripts/launch_floor.py).
// read_only: stream the buffer through L2 with loads instead (L2 is left full of CLEAN lines).
__global__ void __launch_bounds__(256) flush_debug_kernel(uint4* __restrict__ buf, size_t n16, unsigned int v, int read_only) {
  extern __shared__ __align__(16) unsigned char flush_smem[];
  if (threadIdx.x == 0 && v == 0xFFFFFFFFu) flush_smem[0] = 1;  // keep the allocation alive
  const uint4 w = make_uint4(v, v, v, v);
  unsigned int acc = 0u;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (read_only) {
      const uint4 r = __ldcg(buf + i);
      acc ^= r.x ^ r.y ^ r.z ^ r.w;
    } else {
      buf[i] = w;
    }
  }
  if (read_only && acc == 0x9E3779B9u && v == 0xFFFFFFFEu) buf[0] = w;  // (keeps the loads alive)
}

// in [n][6] = (counter x, y, z, w, key lo, key hi) -> out [n][4]: the raw Philox4x32-10 block (known-answer tests)
__global__ void philox_debug_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 r = philox4x32_10(make_uint4(in[6 * i], in[6 * i + 1], in[6 * i + 2], in[6 * i + 3]),
                                make_uint2(in[6 * i + 4], in[6 * i + 5]));
  out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

__global__ void sincos_debug_kernel(const float* __restrict__ th, float* __restrict__ s, float* __restrict__ c, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sincos_heading<false>(th[i], &s[i], &c[i]);
}

#endif  // BNV_ROLLOUT_ONLY

}  // namespace bnv
