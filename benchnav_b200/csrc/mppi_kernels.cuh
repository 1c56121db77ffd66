// sm_100a kernels of the MPPI control iteration (reference: src/planners/local_planners/mppi.py:130-240).
//
//   trav_map_kernel     risk map -> tau = 1 - clamp(risk,0,1), padded pitch          (traversability_model.py:71-72)
//   noise_kernel        Philox4x32-10 + Box-Muller -> sigma-scaled noise [Kl,T,2]     (mppi.py:149-151)
//   rollout_kernel      clamp, T-step unicycle rollout, costs, per-CTA softmax partial, (mppi.py:152-199)
//                       last CTA: grid merge, weights, optimal rollout                (mppi.py:193-217)
//   finalize_kernel     multi-GPU: merge gathered shard partials, weights, optimal rollout
//   top-n kernels       radix select + sort + gather                                  (mppi.py:221-240)
//
// Work decomposition of rollout_kernel: one thread = one sample, state and running cost in registers;
// one warp = 32 consecutive samples with its own noise slab (1-D bulk copy global->shared, own mbarrier)
// and its own recorded-state slab (shared->global bulk store), so warps never synchronise inside the
// T-loop; one CTA = kWarps warps sharing the traversability window staged by a 2-D TMA load.
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

struct alignas(64) EngineParams {
  CUtensorMap tau_map;  // 2-D tiled descriptor over the padded tau map, box = patch_w x patch_h
  const float* tau;     // [G][pitch]
  int G, pitch;
  GridGeom geom;
  Bounds bounds;
  float goal_x, goal_y, thr;
  float lambda, icov0, icov1;  // temperature, 1/sigma^2 (diag of mppi.py:95 inverse covariance)
  int Kl, T;                   // shard-local samples, horizon
  int patch_w, patch_h, rho;   // staged window size (cells) and reach radius
  int use_patch;               // 0: window does not fit shared memory -> look up in the global map (L2)
  int warps;                   // warps per rollout CTA
  int world;                   // number of sample shards
  int record;                  // keep recorded states
  int noise_bulk_ok, rec_bulk_ok;  // pointers 16 B aligned -> bulk copies allowed
  const float* state;   // [3]
  const float* noise;   // [Kl][T][2]
  float* u_prev;        // [T][2]   mean sequence (read at start, replaced by u* at the end)
  float* rec;           // [Kl][T+1][3]
  float* costs;         // [Kl]
  float* weights;       // [Kl]
  float* part_ms;       // [nCTA][2]   per-CTA (max score, sum exp)
  float* part_u;        // [nCTA][2T]  per-CTA sum exp * v
  float* shard_partial; // [2+2T]      (m, s, U) of this shard
  unsigned int* ticket;
  float* u_out;         // [T][2]
  float* opt_rec;       // [T+1][3]
};

// Shared-memory carve-up of the rollout kernel (identical on host and device).
struct RolloutSmem {
  int off_patch, off_noise, off_rec, off_uprev, off_e, off_warpu, off_red, total;
};
__host__ __device__ inline RolloutSmem rollout_smem_layout(int T, int warps, int patch_w, int patch_h, int use_patch,
                                                           int record) {
  RolloutSmem s;
  int spb = warps * 32;
  int off = 128;  // [0,128): mbarriers (1 patch + kMaxWarps noise) and the last-CTA flag
  s.off_patch = off;
  off += use_patch ? ((patch_w * patch_h * 4 + 127) / 128) * 128 : 0;
  s.off_noise = off;
  off += ((spb * 2 * T * 4 + 127) / 128) * 128;
  s.off_rec = off;
  off += record ? ((spb * 3 * (T + 1) * 4 + 127) / 128) * 128 : 0;
  s.off_uprev = off;
  off += ((2 * T * 4 + 15) / 16) * 16;
  s.off_e = off;
  off += spb * 4;
  s.off_warpu = off;
  off += ((kMaxWarps * 2 * T * 4 + 15) / 16) * 16;
  s.off_red = off;
  off += 64 * 4;
  s.total = off;
  return s;
}

// --------------------------------------------------------------------------------------------- tau map
__global__ void trav_map_kernel(const float* __restrict__ risk, int risk_pitch, float* __restrict__ tau, int pitch,
                                int G) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= pitch) return;
  float v = 0.0f;
  if (x < G) {
    float r = risk[static_cast<size_t>(y) * risk_pitch + x];
    // torch.clamp propagates NaN; fminf/fmaxf would not
    float c = (r != r) ? r : fminf(fmaxf(r, 0.0f), 1.0f);
    v = __fsub_rn(1.0f, c);
  }
  tau[static_cast<size_t>(y) * pitch + x] = v;
}

// --------------------------------------------------------------------------------------------- noise
// One thread = one (sample, step pair): 4 normals = noise[k][2p..2p+1][0..1].
__global__ void __launch_bounds__(256) noise_kernel(float* __restrict__ noise, int Kl, int T, int k_offset,
                                                    uint32_t seed_lo, uint32_t seed_hi, uint32_t iter_lo,
                                                    uint32_t iter_hi, float sigma0, float sigma1) {
  int pairs = (T + 1) >> 1;
  long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(Kl) * pairs) return;
  int k = static_cast<int>(gid / pairs);
  int p = static_cast<int>(gid - static_cast<long long>(k) * pairs);
  uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(k + k_offset), static_cast<uint32_t>(p), iter_lo, iter_hi),
                          make_uint2(seed_lo, seed_hi));
  float2 a = box_muller(r.x, r.y);
  float2 b = box_muller(r.z, r.w);
  float* dst = noise + (static_cast<size_t>(k) * T + 2 * p) * 2;
  if (2 * p + 1 < T) {
    if ((T & 1) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(sigma0 * a.x, sigma1 * a.y, sigma0 * b.x, sigma1 * b.y);
    } else {
      *reinterpret_cast<float2*>(dst) = make_float2(sigma0 * a.x, sigma1 * a.y);
      *reinterpret_cast<float2*>(dst + 2) = make_float2(sigma0 * b.x, sigma1 * b.y);
    }
  } else {
    *reinterpret_cast<float2*>(dst) = make_float2(sigma0 * a.x, sigma1 * a.y);
  }
}

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
  int cx = min(max(cell_coord(sx, P.geom.x_min, P.geom), 0), P.G - 1);
  int cy = min(max(cell_coord(sy, P.geom.y_min, P.geom), 0), P.G - 1);
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

// Score of a cost: x = -c / lambda (mppi.py:193), true division.
__device__ __forceinline__ float score_of(float cost, float lambda) { return __fdiv_rn(-cost, lambda); }

// Last phase of an iteration, run by ONE CTA once (M, S, U) over all samples are known:
// u* = U / S (mppi.py:196-199), next mean sequence (mppi.py:217), weights (mppi.py:193) and the batch-1
// optimal rollout (mppi.py:202-214).  `u_s` (shared, 2T floats) holds U on entry and u* on exit.
__device__ void finish_iteration(const EngineParams& P, const TauWindow& win, float M, float S, float* u_s,
                                 float sx, float sy, float sth) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int T = P.T;
  for (int c = tid; c < 2 * T; c += nthr) {
    float u = __fdiv_rn(u_s[c], S);
    u_s[c] = u;
    P.u_out[c] = u;
    P.u_prev[c] = u;
  }
  __syncthreads();
  if (tid == 0) {
    float x = sx, y = sy, th = sth;
    float tau = lookup_tau(win, P.geom, x, y);
    for (int t = 0; t < T; ++t) {
      float xr, yr, thr;
      unicycle_step(P.geom, P.bounds, tau, u_s[2 * t], u_s[2 * t + 1], x, y, th, xr, yr, thr);
      P.opt_rec[3 * t + 0] = xr;
      P.opt_rec[3 * t + 1] = yr;
      P.opt_rec[3 * t + 2] = thr;
      tau = lookup_tau(win, P.geom, x, y);
    }
    P.opt_rec[3 * T + 0] = x;
    P.opt_rec[3 * T + 1] = y;
    P.opt_rec[3 * T + 2] = th;
  }
  // weights w_k = exp(x_k - M) / S over the shard (mppi.py:193).  With more than one warp the serial optimal
  // rollout keeps warp 0 busy and the other warps normalise underneath it; loads are batched for MLP.
  const bool split = nthr > 32;
  if (split && tid < 32) return;
  if (!split) __syncwarp();
  const int wt = split ? tid - 32 : tid, wn = split ? nthr - 32 : nthr;
  const float inv_s = __fdiv_rn(1.0f, S);
  constexpr int kBatch = 8;
  for (int base = 0; base < P.Kl; base += wn * kBatch) {
    float c[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      int k = base + j * wn + wt;
      c[j] = (k < P.Kl) ? __ldcg(P.costs + k) : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      int k = base + j * wn + wt;
      if (k < P.Kl) P.weights[k] = expf(score_of(c[j], P.lambda) - M) * inv_s;
    }
  }
}

// --------------------------------------------------------------------------------------------- rollout
__global__ void __launch_bounds__(kMaxWarps * 32, 1) rollout_kernel(const __grid_constant__ EngineParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = P.T;
  const int spb = P.warps * 32;
  const RolloutSmem L = rollout_smem_layout(T, P.warps, P.patch_w, P.patch_h, P.use_patch, P.record);
  uint64_t* bar_patch = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar_noise = reinterpret_cast<uint64_t*>(smem) + 1;  // [kMaxWarps]
  int* last_flag = reinterpret_cast<int*>(smem + 64);
  float* patch_s = reinterpret_cast<float*>(smem + L.off_patch);
  float* noise_s = reinterpret_cast<float*>(smem + L.off_noise);
  float* rec_s = reinterpret_cast<float*>(smem + L.off_rec);
  float* uprev_s = reinterpret_cast<float*>(smem + L.off_uprev);
  float* e_s = reinterpret_cast<float*>(smem + L.off_e);
  float* warpu_s = reinterpret_cast<float*>(smem + L.off_warpu);
  float* red_s = reinterpret_cast<float*>(smem + L.off_red);

  const int cta_first = blockIdx.x * spb;
  const int warp_first = cta_first + warp * 32;
  const int warp_rows = max(0, min(32, P.Kl - warp_first));
  const int k = warp_first + lane;
  const bool valid = lane < warp_rows;

  if (tid == 0) {
    mbar_init(bar_patch, 1);
    for (int w = 0; w < kMaxWarps; ++w) mbar_init(bar_noise + w, 1);
    fence_mbar_init();
    if (P.use_patch) prefetch_tensormap(&P.tau_map);
  }
  __syncthreads();

  // ---- stage this warp's noise slab: rows [warp_first, warp_first+warp_rows) x 2T floats, contiguous in HBM
  float* nz_w = noise_s + warp * 32 * 2 * T;
  const float* nz_g = P.noise + static_cast<size_t>(warp_first) * 2 * T;
  const uint32_t nz_bytes = static_cast<uint32_t>(warp_rows) * 2u * T * 4u;
  const bool nz_bulk = P.noise_bulk_ok && ((nz_bytes & 15u) == 0u) && warp_rows > 0;
  if (nz_bulk) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_noise + warp, nz_bytes);
      bulk_load_g2s(nz_w, nz_g, nz_bytes, bar_noise + warp);
    }
  } else {
    for (int i = lane; i < warp_rows * 2 * T; i += 32) nz_w[i] = nz_g[i];
  }

  // ---- state, window geometry, traversability window via TMA
  const float sx = P.state[0], sy = P.state[1], sth = P.state[2];
  const WindowGeom wg = window_for_state(P, sx, sy);
  TauWindow win;
  if (P.use_patch) {
    if (warp == 0) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_patch, static_cast<uint32_t>(P.patch_w * P.patch_h * 4));
        tma_load_2d(patch_s, &P.tau_map, wg.ox, wg.oy, bar_patch);
      }
    }
    win.base = patch_s - (wg.oy * P.patch_w + wg.ox);
    win.pitch = P.patch_w;
  } else {
    win.base = P.tau;
    win.pitch = P.pitch;
  }
  win.lo_x = wg.lo_x;
  win.hi_x = wg.hi_x;
  win.lo_y = wg.lo_y;
  win.hi_y = wg.hi_y;

  for (int c = tid; c < 2 * T; c += blockDim.x) uprev_s[c] = P.u_prev[c];
  __syncthreads();
  if (P.use_patch) mbar_wait(bar_patch, 0);
  if (nz_bulk) mbar_wait(bar_noise + warp, 0);
  else __syncwarp();

  // ---- T-step rollout, one sample per thread
  float cost = FLT_MAX;
  const float* nrow = nz_w + lane * 2 * T;
  float* rrow = rec_s + (warp * 32 + lane) * 3 * (T + 1);
  if (valid) {
    float x = sx, y = sy, th = sth;
    float tau = lookup_tau(win, P.geom, x, y);
    float stage_sum = 0.0f, act_sum = 0.0f;
    for (int t = 0; t < T; ++t) {
      const float2 n = *reinterpret_cast<const float2*>(nrow + 2 * t);
      const float2 up = *reinterpret_cast<const float2*>(uprev_s + 2 * t);
      const float v0 = clampf(__fadd_rn(up.x, n.x), P.bounds.u_min0, P.bounds.u_max0);  // mppi.py:152-157
      const float v1 = clampf(__fadd_rn(up.y, n.y), P.bounds.u_min1, P.bounds.u_max1);
      float xr, yr, thr;
      unicycle_step(P.geom, P.bounds, tau, v0, v1, x, y, th, xr, yr, thr);
      if (P.record) {
        rrow[3 * t + 0] = xr;
        rrow[3 * t + 1] = yr;
        rrow[3 * t + 2] = thr;
      }
      // one lookup serves the stage cost of the recorded (raw) position and the next dynamics step
      tau = lookup_tau(win, P.geom, x, y);
      stage_sum = __fadd_rn(stage_sum, goal_and_stuck_cost(xr, yr, P.goal_x, P.goal_y, tau, P.thr));
      const float act = __fadd_rn(__fmul_rn(__fmul_rn(up.x, P.icov0), v0), __fmul_rn(__fmul_rn(up.y, P.icov1), v1));
      act_sum = __fadd_rn(act_sum, __fmul_rn(P.lambda, act));  // mppi.py:178-182, :189
    }
    if (P.record) {
      rrow[3 * T + 0] = x;
      rrow[3 * T + 1] = y;
      rrow[3 * T + 2] = th;
    }
    const float terminal = goal_and_stuck_cost(x, y, P.goal_x, P.goal_y, tau, P.thr);  // mppi.py:184
    cost = __fadd_rn(__fadd_rn(stage_sum, terminal), act_sum);                         // mppi.py:186-190
    P.costs[k] = cost;
  }

  // ---- recorded states: shared -> HBM, one bulk store per warp slab, drains behind the epilogue
  if (P.record && warp_rows > 0) {
    float* rec_g = P.rec + static_cast<size_t>(warp_first) * 3 * (T + 1);
    const uint32_t rec_bytes = static_cast<uint32_t>(warp_rows) * 3u * (T + 1) * 4u;
    float* rec_w = rec_s + warp * 32 * 3 * (T + 1);
    if (P.rec_bulk_ok && (rec_bytes & 15u) == 0u) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        bulk_store_s2g(rec_g, rec_w, rec_bytes);
        bulk_commit();
      }
    } else {
      __syncwarp();
      for (int i = lane; i < warp_rows * 3 * (T + 1); i += 32) rec_g[i] = rec_w[i];
    }
  }

  // ---- per-CTA softmax partial: m = max score, s = sum exp(score - m), U[c] = sum exp(score - m) * v[k][c]
  const float score = valid ? score_of(cost, P.lambda) : -FLT_MAX;
  float wm = warp_max(score);
  if (lane == 0) red_s[warp] = wm;
  __syncthreads();
  float m_cta = red_s[0];
  for (int w = 1; w < P.warps; ++w) m_cta = fmaxf(m_cta, red_s[w]);
  const float e = valid ? expf(score - m_cta) : 0.0f;
  e_s[tid] = e;
  float ws = warp_sum(e);
  if (lane == 0) red_s[8 + warp] = ws;
  __syncwarp();
  // each warp: columns over lanes, its own 32 samples
  for (int c = lane; c < 2 * T; c += 32) {
    const float up = uprev_s[c];
    const float lo = (c & 1) ? P.bounds.u_min1 : P.bounds.u_min0;
    const float hi = (c & 1) ? P.bounds.u_max1 : P.bounds.u_max0;
    float acc = 0.0f;
#pragma unroll 8
    for (int r = 0; r < warp_rows; ++r) {
      const float v = clampf(__fadd_rn(up, nz_w[r * 2 * T + c]), lo, hi);
      acc = fmaf(e_s[warp * 32 + r], v, acc);
    }
    warpu_s[warp * 2 * T + c] = acc;
  }
  __syncthreads();
  float s_cta = 0.0f;
  for (int w = 0; w < P.warps; ++w) s_cta += red_s[8 + w];
  for (int c = tid; c < 2 * T; c += blockDim.x) {
    float acc = 0.0f;
    for (int w = 0; w < P.warps; ++w) acc += warpu_s[w * 2 * T + c];
    P.part_u[static_cast<size_t>(blockIdx.x) * 2 * T + c] = acc;
  }
  if (tid == 0) {
    P.part_ms[2 * blockIdx.x + 0] = m_cta;
    P.part_ms[2 * blockIdx.x + 1] = s_cta;
  }

  // ---- grid-wide merge by the last CTA to finish (atomic ticket)
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned int prev = atomicAdd(P.ticket, 1u);
    *last_flag = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (*last_flag) {
    __threadfence();
    const int nblk = gridDim.x;
    // global max score and sum: M = max_g m_g, S = sum_g exp(m_g - M) s_g
    float lm = -FLT_MAX;
    for (int g = tid; g < nblk; g += blockDim.x) lm = fmaxf(lm, __ldcg(P.part_ms + 2 * g));
    lm = warp_max(lm);
    if (lane == 0) red_s[16 + warp] = lm;
    __syncthreads();
    float M = red_s[16];
    for (int w = 1; w < P.warps; ++w) M = fmaxf(M, red_s[16 + w]);
    float lsum = 0.0f;
    for (int g = tid; g < nblk; g += blockDim.x)
      lsum += expf(__ldcg(P.part_ms + 2 * g) - M) * __ldcg(P.part_ms + 2 * g + 1);
    lsum = warp_sum(lsum);
    if (lane == 0) red_s[24 + warp] = lsum;
    // U[c] = sum_g a_g U_g[c]: warps split g, lanes split columns; all loads independent
    const int ncol = 2 * T;
    for (int c0 = 0; c0 < ncol; c0 += 128) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int g = warp; g < nblk; g += P.warps) {
        const float a = expf(__ldcg(P.part_ms + 2 * g) - M);
        const float* row = P.part_u + static_cast<size_t>(g) * ncol + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = lane + 32 * j;
          if (c0 + c < ncol) acc[j] = fmaf(a, __ldcg(row + c), acc[j]);
        }
      }
      __syncthreads();  // warpu_s reuse across c0 chunks / previous phase
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = lane + 32 * j;
        if (c0 + c < ncol) warpu_s[warp * ncol + c0 + c] = acc[j];
      }
    }
    __syncthreads();
    float S = 0.0f;
    for (int w = 0; w < P.warps; ++w) S += red_s[24 + w];
    for (int c = tid; c < ncol; c += blockDim.x) {
      float acc = 0.0f;
      for (int w = 0; w < P.warps; ++w) acc += warpu_s[w * ncol + c];
      uprev_s[c] = acc;  // U (un-normalised)
    }
    __syncthreads();
    if (tid == 0) *P.ticket = 0u;  // re-arm for the next launch
    if (P.world == 1) {
      finish_iteration(P, win, M, S, uprev_s, sx, sy, sth);
    } else {
      if (tid == 0) {
        P.shard_partial[0] = M;
        P.shard_partial[1] = S;
      }
      for (int c = tid; c < ncol; c += blockDim.x) P.shard_partial[2 + c] = uprev_s[c];
    }
  }
  // shared memory must stay allocated until the bulk stores have read it
  if (P.record && lane == 0) bulk_wait_read_all();
}

// --------------------------------------------------------------------------------------------- finalize (multi-GPU)
// gathered: [world][2+2T] shard partials, identical on every rank -> every rank computes the same u*.
// Block 0 runs finish_iteration for this shard (weights of the local samples, outputs, optimal rollout).
__global__ void __launch_bounds__(kFinalizeThreads) finalize_kernel(const __grid_constant__ EngineParams P,
                                                                    const float* __restrict__ gathered) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int T = P.T, ncol = 2 * T, plen = 2 + 2 * T;
  uint64_t* bar_patch = reinterpret_cast<uint64_t*>(smem);
  float* patch_s = reinterpret_cast<float*>(smem + 128);
  float* u_s = patch_s + (P.use_patch ? ((P.patch_w * P.patch_h + 31) / 32) * 32 : 0);
  if (tid == 0) {
    mbar_init(bar_patch, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const float sx = P.state[0], sy = P.state[1], sth = P.state[2];
  const WindowGeom wg = window_for_state(P, sx, sy);
  TauWindow win;
  if (P.use_patch) {
    if (tid < 32) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_patch, static_cast<uint32_t>(P.patch_w * P.patch_h * 4));
        tma_load_2d(patch_s, &P.tau_map, wg.ox, wg.oy, bar_patch);
      }
    }
    win.base = patch_s - (wg.oy * P.patch_w + wg.ox);
    win.pitch = P.patch_w;
  } else {
    win.base = P.tau;
    win.pitch = P.pitch;
  }
  win.lo_x = wg.lo_x;
  win.hi_x = wg.hi_x;
  win.lo_y = wg.lo_y;
  win.hi_y = wg.hi_y;

  float M = -FLT_MAX;
  for (int g = 0; g < P.world; ++g) M = fmaxf(M, gathered[g * plen]);
  float S = 0.0f;
  for (int g = 0; g < P.world; ++g) S += expf(gathered[g * plen] - M) * gathered[g * plen + 1];
  for (int c = tid; c < ncol; c += blockDim.x) {
    float acc = 0.0f;
    for (int g = 0; g < P.world; ++g) acc = fmaf(expf(gathered[g * plen] - M), gathered[g * plen + 2 + c], acc);
    u_s[c] = acc;
  }
  __syncthreads();
  if (P.use_patch) mbar_wait(bar_patch, 0);
  finish_iteration(P, win, M, S, u_s, sx, sy, sth);
}

// --------------------------------------------------------------------------------------------- top-n
// MPPI.get_top_samples (mppi.py:221-240).  One CTA: 4-pass byte-wise radix select of the n-th largest
// weight (weights are >= 0, so their bit patterns order like unsigned integers), compaction of the n
// winners, bitonic sort (descending weight, ascending index among ties) and output.  `pairs` is a scratch
// of n_pad (power of two >= n) 64-bit words in shared memory (n_pad <= kTopnSmemPairs) or global memory.
constexpr int kTopnThreads = 1024;
constexpr int kTopnSmemPairs = 16384;

__global__ void __launch_bounds__(kTopnThreads) topn_select_kernel(const float* __restrict__ weights, int K, int n,
                                                                   int n_pad, unsigned long long* pairs_global,
                                                                   float* __restrict__ out_w, int* __restrict__ out_idx) {
  extern __shared__ __align__(128) unsigned char smem[];
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
  unsigned int mask = 0u;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const unsigned int prefix = sel_prefix;
    for (int i = tid; i < K; i += blockDim.x) {
      unsigned int key = __float_as_uint(weights[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFFu], 1u);
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
  for (int i = tid; i < K; i += blockDim.x) {
    unsigned int key = __float_as_uint(weights[i]);
    unsigned long long packed = (static_cast<unsigned long long>(key) << 32) | static_cast<unsigned int>(~i);
    if (key > thr_key) {
      unsigned int slot = atomicAdd(&n_gt, 1u);
      pairs[slot] = packed;
    } else if (key == thr_key) {
      unsigned int slot = atomicAdd(&n_eq, 1u);
      if (slot < take_eq) pairs[first_eq + slot] = packed;
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
}

// out[i][:] = rec[idx[i]][:], row = 3 (T+1) floats.
__global__ void gather_rows_kernel(const float* __restrict__ rec, const int* __restrict__ idx, int row_len,
                                   float* __restrict__ out) {
  const float* src = rec + static_cast<size_t>(idx[blockIdx.x]) * row_len;
  float* dst = out + static_cast<size_t>(blockIdx.x) * row_len;
  for (int i = threadIdx.x; i < row_len; i += blockDim.x) dst[i] = src[i];
}

// --------------------------------------------------------------------------------------------- debug
__global__ void sincos_debug_kernel(const float* __restrict__ th, float* __restrict__ s, float* __restrict__ c, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sincos_heading(th[i], &s[i], &c[i]);
}

}  // namespace bnv
