// Device math of the MPPI hot path: cell lookup, in-range sin/cos, angle wrap, Philox noise.
// Everything that feeds the recorded states is written with explicit round-to-nearest intrinsics so that
// nvcc cannot contract the reference's mul/mul/mul/add sequences into FMAs (DESIGN.md "Parity arithmetic").
//
// Two flavours of the per-step math exist: a GENERAL one (any heading, any angular step; has slow-path
// branches) used for the first step of every rollout, and a FAST, branch-free one used for steps 1..T-1,
// valid because after one step the heading lies in [-pi, pi) and moves by less than pi per step
// (host-checked: dt * max|omega| < pi).  Both produce bit-identical results on the fast domain.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bnv {

constexpr float kPi = 3.14159274101257324f;     // float32(torch.pi)
constexpr float kTwoPi = 6.28318548202514648f;  // float32(2 * torch.pi)
constexpr float kStuckPenalty = 1e4f;           // objectives.py:53

// floor() on the FMA pipe: for |q| < 2^22, q + 1.5*2^23 rounded towards -inf has floor(q) in its low mantissa
// bits, i.e. __float_as_int(result) == kMagicBits + floor(q).  Replaces the F2I conversion on the per-step chain.
constexpr float kMagicFloat = 12582912.0f;  // 1.5 * 2^23
constexpr int kMagicBits = 0x4B400000;      // its bit pattern

// Keep a loop-invariant value in a register: without this ptxas re-loads kernel parameters from the constant
// bank inside the T-loop (each reload sits on the dependency chain of a warp that runs alone on its scheduler).
__device__ __forceinline__ void pin(float& v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void pin(int& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void pin(uint32_t& v) { asm volatile("" : "+r"(v)); }

struct GridGeom {
  float x_min, y_min, x_max, y_max;
  float res, inv_res;
  int res_pow2;  // resolution is a power of two: (p - p_min) * inv_res == (p - p_min) / res exactly
  int fast_grid; // ... and the map origin is (0, 0) (every GridMap built by grid_map.py:42-50): p - p_min == p, so the
                 // biased cell coordinate is ONE round-down FMA, p * inv_res + magic (the product is exact) -- the
                 // rollout kernels' kPow2 instantiations
};
struct Bounds {
  float u_min0, u_min1, u_max0, u_max1, dt;
};

// Loop-invariant operands of one rollout step, all register-resident.
struct StepConsts {
  float x_min, y_min, x_max, y_max, res, inv_res, dt;
  float gx, gy, thr;
  float u_min0, u_min1, u_max0, u_max1;
  int lo_x, hi_x, lo_y, hi_y;  // clamp range of the cell index = map range intersected with the staged window
  int pitch;                    // row pitch (elements) of the table being indexed
  uint32_t win_addr;            // kPatch: shared-memory byte address of window cell (0,0) in map coordinates
  const float* map;             // !kPatch: global traversability map
  // "magic floor" variants (see lookup_tau): indices carry the bias kMagicBits, folded into bounds and base
  int mlo_x, mhi_x, mlo_y, mhi_y;
  uint32_t mwin_addr;
  // elem_bytes: 4 (one traversability value per cell) or 8 (interleaved slip mean/std per cell, stochastic mode)
  __device__ __forceinline__ void finish(uint32_t elem_bytes = 4u) {
    mlo_x = lo_x + kMagicBits; mhi_x = hi_x + kMagicBits; mlo_y = lo_y + kMagicBits; mhi_y = hi_y + kMagicBits;
    mwin_addr = win_addr - elem_bytes * (static_cast<uint32_t>(kMagicBits) * static_cast<uint32_t>(pitch) +
                                         static_cast<uint32_t>(kMagicBits));
  }
  __device__ __forceinline__ void pin_all() {
    pin(x_min); pin(y_min); pin(x_max); pin(y_max); pin(res); pin(inv_res); pin(dt);
    pin(gx); pin(gy); pin(thr); pin(u_min0); pin(u_min1); pin(u_max0); pin(u_max1);
    pin(lo_x); pin(hi_x); pin(lo_y); pin(hi_y); pin(pitch); pin(win_addr);
    pin(mlo_x); pin(mhi_x); pin(mlo_y); pin(mhi_y); pin(mwin_addr);
  }
};

// ---------------------------------------------------------------------------------------------
// Traversability lookup.  Reference: GridMap.get_grid_indices_from_positions (grid_map.py:195-209):
//   idx = clamp(int(floor((p - p_min) / r)), 0, G-1), value = map[iy, ix]       (grid_map.py:167)
// The engine looks up tau = 1 - clamp(risk,0,1) (traversability_model.py:71-72) precomputed per cell.
// Every position a rollout can reach lies inside the staged window (DESIGN.md "Reach bound"), so clamping
// to [lo, hi] gives the same cell as the reference's clamp to [0, G-1] and is memory-safe regardless.
// ---------------------------------------------------------------------------------------------
template <bool kPow2>
__device__ __forceinline__ int cell_coord(float p, float p_min, float res, float inv_res) {
  float d = __fsub_rn(p, p_min);
  float q = kPow2 ? __fmul_rn(d, inv_res) : __fdiv_rn(d, res);  // true division (CPU ATen semantics)
  return __float2int_rd(q);  // floor + convert in one instruction (saturating; NaN -> 0)
}
__device__ __forceinline__ int cell_coord_rt(float p, float p_min, const GridGeom& g) {
  return g.res_pow2 ? cell_coord<true>(p, p_min, g.res, g.inv_res) : cell_coord<false>(p, p_min, g.res, g.inv_res);
}

// kMagic: the position is known to lie inside the map limits (every state after the first step is clamped to
// them), so floor() can use the magic-constant add; the bias stays in the index and is folded into the clamp
// bounds and the table base.  Same cell as the F2I path for every in-range position.
//
// cell_ref(): where the cell's entry lives -- the shared-memory byte address (kPatch) or the element index in the
// global table -- for entries of kElem bytes; two positions are in the same cell iff their cell_ref is equal.
// kPow2 here means GridGeom::fast_grid (power-of-two resolution AND zero origin); otherwise true division.
template <bool kPatch, bool kPow2, bool kMagic, uint32_t kElem>
__device__ __forceinline__ uint32_t cell_ref(const StepConsts& c, float x, float y) {
  if (kMagic) {
    float bx, by;  // floor((p - p_min) / res) + magic, rounded down
    if (kPow2) {
      bx = __fmaf_rd(x, c.inv_res, kMagicFloat);
      by = __fmaf_rd(y, c.inv_res, kMagicFloat);
    } else {
      bx = __fadd_rd(__fdiv_rn(__fsub_rn(x, c.x_min), c.res), kMagicFloat);
      by = __fadd_rd(__fdiv_rn(__fsub_rn(y, c.y_min), c.res), kMagicFloat);
    }
    int ix = min(max(__float_as_int(bx), c.mlo_x), c.mhi_x);
    int iy = min(max(__float_as_int(by), c.mlo_y), c.mhi_y);
    if (kPatch)
      return c.mwin_addr + kElem * (static_cast<uint32_t>(iy) * static_cast<uint32_t>(c.pitch) + static_cast<uint32_t>(ix));
    return static_cast<uint32_t>((iy - kMagicBits) * c.pitch + (ix - kMagicBits));
  }
  // general first lookup (position possibly outside the map): x - 0 == x, so kPow2 may keep the subtraction
  int ix = min(max(cell_coord<kPow2>(x, c.x_min, c.res, c.inv_res), c.lo_x), c.hi_x);
  int iy = min(max(cell_coord<kPow2>(y, c.y_min, c.res, c.inv_res), c.lo_y), c.hi_y);
  const uint32_t idx = static_cast<uint32_t>(iy * c.pitch + ix);
  return kPatch ? c.win_addr + kElem * idx : idx;
}
template <bool kPatch>
__device__ __forceinline__ float load_tau(const StepConsts& c, uint32_t ref) {
  if (kPatch) {
    float t;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(ref));
    return t;
  }
  return __ldg(c.map + ref);
}
template <bool kPatch>
__device__ __forceinline__ float2 load_slip(const StepConsts& c, uint32_t ref) {
  if (kPatch) {
    float2 t;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "r"(ref));
    return t;
  }
  return __ldg(reinterpret_cast<const float2*>(c.map) + ref);
}
template <bool kPatch, bool kPow2, bool kMagic>
__device__ __forceinline__ float lookup_tau(const StepConsts& c, float x, float y) {
  return load_tau<kPatch>(c, cell_ref<kPatch, kPow2, kMagic, 4u>(c, x, y));
}

// Stochastic-slip mode (BASELINE config 4; observation-mode lookup, traversability_model.py:65-69 +
// grid_map.py:169-178): the table holds (mean, std) of the cell's slip distribution interleaved, 8 bytes per cell.
// Same cell index as lookup_tau; returns (mean, std).
template <bool kPatch, bool kPow2, bool kMagic>
__device__ __forceinline__ float2 lookup_slip(const StepConsts& c, float x, float y) {
  return load_slip<kPatch>(c, cell_ref<kPatch, kPow2, kMagic, 8u>(c, x, y));
}

// 1 - clamp(sample, 0, 1) with sample = xi * std + mean: Normal(mean, std).sample() is ATen's
// normal_(0,1).mul_(std).add_(mean) (two roundings, no FMA); torch.clamp propagates NaN.
__device__ __forceinline__ float slip_to_trav(float2 ms, float xi) {
  float smp = __fadd_rn(__fmul_rn(xi, ms.y), ms.x);
  float c = (smp != smp) ? smp : fminf(fmaxf(smp, 0.0f), 1.0f);
  return __fsub_rn(1.0f, c);
}

// ---------------------------------------------------------------------------------------------
// sin/cos for the heading: 3-term Cody-Waite reduction by pi/2 (exact for the small quotients that occur)
// and degree-7/8 minimax polynomials; <= 2 ulp on |x| <= 64 (tests/test_parity_gpu.py::test_sincos_accuracy),
// the same class as torch's CPU (Sleef u10) and CUDA kernels.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sincos_reduced(float x, float* sn, float* cs) {
  // q = round(x * 2/pi) by the add-magic-constant trick (stays on the FMA pipe; no FRND/F2I conversions):
  // the low mantissa bits of t hold the integer quotient in two's complement.
  float t = fmaf(x, 0.636619747f, 12582912.0f);
  int n = __float_as_int(t);
  float q = t - 12582912.0f;
  float r = fmaf(q, -1.57079601287841796875f, x);  // pi/2 split hi + mid: the product with hi is exact for the
  r = fmaf(q, -3.1391647326017846e-7f, r);         // small quotients that occur (|q| <= 41); residual 5.4e-15 |q|
  float z = r * r;
  float ps = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);  // sin(r), |r| <= pi/4
  ps = fmaf(ps, z, -1.6666654611e-1f);
  float s = fmaf(ps * z, r, r);
  float pc = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);  // cos(r), |r| <= pi/4
  pc = fmaf(pc, z, 4.166664568298827e-2f);
  float c = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
  float s_out = (n & 1) ? c : s;
  float c_out = (n & 1) ? s : c;
  // sign flips as integer XORs on the sign bit (branch-free)
  uint32_t s_sign = (static_cast<uint32_t>(n) & 2u) << 30;
  uint32_t c_sign = (static_cast<uint32_t>(n + 1) & 2u) << 30;
  *sn = __uint_as_float(__float_as_uint(s_out) ^ s_sign);
  *cs = __uint_as_float(__float_as_uint(c_out) ^ c_sign);
}

template <bool kFast>
__device__ __forceinline__ void sincos_heading(float x, float* sn, float* cs) {
  if (!kFast && !(fabsf(x) <= 64.0f)) {  // also catches NaN/inf
    sincosf(x, sn, cs);
    return;
  }
  sincos_reduced(x, sn, cs);
}

// ---------------------------------------------------------------------------------------------
// Angle wrap: (theta + pi) % (2 pi) - pi with torch.remainder semantics (fmod, then add the divisor
// when the result is non-zero and negative) -- robot_model.py:90.  For a = theta + pi in (-2pi, 4pi)
// fmod is a select between a, a - 2pi (exact, Sterbenz) and a itself: branch-free.
// ---------------------------------------------------------------------------------------------
template <bool kFast>
__device__ __forceinline__ float wrap_heading(float theta_raw) {
  float a = __fadd_rn(theta_raw, kPi);
  float m = a;
  if (kFast || (a > -kTwoPi && a < 2.0f * kTwoPi)) {
    m = (a >= kTwoPi) ? __fsub_rn(a, kTwoPi) : m;
    m = (a < 0.0f) ? __fadd_rn(a, kTwoPi) : m;
  } else {
    m = fmodf(a, kTwoPi);
    if (m != 0.0f && m < 0.0f) m = __fadd_rn(m, kTwoPi);
  }
  return __fsub_rn(m, kPi);
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// One unicycle step (UnicycleModel.transit, robot_model.py:75-95) for controls already clamped to the action
// bounds.  (x, y, th) is the clamped/wrapped state, tau its traversability; writes the raw successor (what the
// reference's in-place `+=` leaves in the input slot) and advances (x, y, th) to the clamped/wrapped successor.
template <bool kFast>
__device__ __forceinline__ void unicycle_step(const StepConsts& c, float tau, float v0, float v1, float& x, float& y,
                                              float& th, float& xr, float& yr, float& thr) {
  float sn, cs;
  sincos_heading<kFast>(th, &sn, &cs);
  float tv = __fmul_rn(tau, v0);
  xr = __fadd_rn(x, __fmul_rn(__fmul_rn(tv, cs), c.dt));  // x += trav * v * cos(theta) * dt
  yr = __fadd_rn(y, __fmul_rn(__fmul_rn(tv, sn), c.dt));
  thr = __fadd_rn(th, __fmul_rn(__fmul_rn(tau, v1), c.dt));
  x = clampf(xr, c.x_min, c.x_max);
  y = clampf(yr, c.y_min, c.y_max);
  th = wrap_heading<kFast>(thr);
}

// Stage/terminal cost term (objectives.py:46-53): ||p - goal|| + 1e4 * [tau <= thr].  sqrt.approx has a
// maximum relative error of 2^-23 (PTX ISA), far inside the cost tolerance, and no slow-path branch.
__device__ __forceinline__ float goal_and_stuck_cost_at(const StepConsts& c, float gx, float gy, float px, float py,
                                                        float tau) {
  float dx = __fsub_rn(px, gx), dy = __fsub_rn(py, gy);
  float d2 = fmaf(dx, dx, dy * dy);
  float d;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(d2));  // ftz: squared distances below 1e-38 m^2 read as 0
  return __fadd_rn(d, (tau <= c.thr) ? kStuckPenalty : 0.0f);
}
__device__ __forceinline__ float goal_and_stuck_cost(const StepConsts& c, float px, float py, float tau) {
  return goal_and_stuck_cost_at(c, c.gx, c.gy, px, py, tau);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator + Box-Muller.  Counter = (global sample, step pair, iteration),
// key = seed, so a sample's noise does not depend on how samples are sharded over GPUs.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// Two independent standard normals from two 32-bit words (Box-Muller on the SFU): u1 = (a + 0.5) 2^-32 in (0,1),
// r = sqrt(-2 ln u1), angle = 2 pi ((b + 0.5) 2^-32 - 0.5) in (-pi, pi).  lg2/sqrt/sin/cos are the approximate
// MUFU forms (abs. error ~1e-6 on a unit normal, far below the noise's own scale; moments checked by
// tests/test_parity_gpu.py::test_philox_noise_statistics_and_determinism).  This function DEFINES the engine's
// noise stream: the in-loop draw and the stand-alone noise kernel both call it.
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  float u1 = fmaf(static_cast<float>(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // a*2^-32 + 2^-33
  float ang = fmaf(static_cast<float>(b), 1.4629180792671596e-9f, -3.1415926535897931f + 7.3145903963357981e-10f);
  float l2, r, sn, cs;
  asm("lg2.approx.f32 %0, %1;" : "=f"(l2) : "f"(u1));
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(l2 * -1.3862943611198906f));  // -2 ln 2 * log2(u1)
  asm("sin.approx.f32 %0, %1;" : "=f"(sn) : "f"(ang));
  asm("cos.approx.f32 %0, %1;" : "=f"(cs) : "f"(ang));
  return make_float2(r * cs, r * sn);
}

// Sigma-scaled noise of one sample for the step pair (2p, 2p+1): (n[2p][0], n[2p][1], n[2p+1][0], n[2p+1][1]).
__device__ __forceinline__ float4 noise_pair(uint32_t sample, uint32_t pair, uint32_t iter_lo, uint32_t iter_hi,
                                             uint2 key, float sigma0, float sigma1) {
  const uint4 r = philox4x32_10(make_uint4(sample, pair, iter_lo, iter_hi), key);
  const float2 a = box_muller(r.x, r.y);
  const float2 b = box_muller(r.z, r.w);
  return make_float4(sigma0 * a.x, sigma1 * a.y, sigma0 * b.x, sigma1 * b.y);
}

// Standard normals for the stochastic-slip lookups of one sample: stream `pair | 0x80000000` of the same
// generator (the control noise uses pair < 2^31).  For step pair p: (transit 2p, stage 2p, transit 2p+1, stage 2p+1);
// pair 0x7FFFFFFF: .x = terminal lookup.
constexpr uint32_t kXiStream = 0x80000000u;
constexpr uint32_t kXiTerminalPair = 0x7FFFFFFFu;
constexpr uint32_t kOptimalSample = 0xFFFFFFFFu;  // "sample" id of the batch-1 optimal rollout's lookups
__device__ __forceinline__ float4 xi_quad(uint32_t sample, uint32_t pair, uint32_t iter_lo, uint32_t iter_hi, uint2 key) {
  const uint4 r = philox4x32_10(make_uint4(sample, pair | kXiStream, iter_lo, iter_hi), key);
  const float2 a = box_muller(r.x, r.y);
  const float2 b = box_muller(r.z, r.w);
  return make_float4(a.x, a.y, b.x, b.y);
}

}  // namespace bnv
