// Device math of the MPPI hot path: cell lookup, in-range sin/cos, angle wrap, Philox noise.
// Everything that feeds the recorded states is written with explicit round-to-nearest intrinsics so that
// nvcc cannot contract the reference's mul/mul/mul/add sequences into FMAs (DESIGN.md "Parity arithmetic").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bnv {

constexpr float kPi = 3.14159274101257324f;     // float32(torch.pi)
constexpr float kTwoPi = 6.28318548202514648f;  // float32(2 * torch.pi)
constexpr float kStuckPenalty = 1e4f;           // objectives.py:53

// ---------------------------------------------------------------------------------------------
// Traversability lookup.  Reference: GridMap.get_grid_indices_from_positions (grid_map.py:195-209):
//   idx = clamp(int(floor((p - p_min) / r)), 0, G-1), value = map[iy, ix]       (grid_map.py:167)
// The engine looks up tau = 1 - clamp(risk,0,1) (traversability_model.py:71-72) precomputed per cell.
// `base` is pre-offset so that base[iy * pitch + ix] is valid for ix in [lo_x, hi_x], iy in [lo_y, hi_y];
// [lo, hi] is the intersection of the map [0, G-1] with the staged window.  Every position the rollout
// can reach lies inside the window (DESIGN.md "Reach bound"), so clamping to [lo, hi] gives the same
// cell as the reference's clamp to [0, G-1] and is memory-safe regardless.
// ---------------------------------------------------------------------------------------------
struct TauWindow {
  const float* base;
  int pitch;
  int lo_x, hi_x, lo_y, hi_y;
};

struct GridGeom {
  float x_min, y_min, x_max, y_max;
  float res, inv_res;
  int res_pow2;  // resolution is a power of two: (p - p_min) * inv_res == (p - p_min) / res exactly
};

__device__ __forceinline__ int cell_coord(float p, float p_min, const GridGeom& g) {
  float d = __fsub_rn(p, p_min);
  float q = g.res_pow2 ? __fmul_rn(d, g.inv_res) : __fdiv_rn(d, g.res);  // true division (CPU ATen semantics)
  return __float2int_rd(q);  // floor + convert in one instruction (saturating; NaN -> 0)
}

__device__ __forceinline__ float lookup_tau(const TauWindow& w, const GridGeom& g, float x, float y) {
  int ix = min(max(cell_coord(x, g.x_min, g), w.lo_x), w.hi_x);
  int iy = min(max(cell_coord(y, g.y_min, g), w.lo_y), w.hi_y);
  return w.base[iy * w.pitch + ix];
}

// ---------------------------------------------------------------------------------------------
// sin/cos for the heading.  The heading is wrapped into [-pi, pi) after every step and moves by at most
// dt*|omega|max per step, so a 3-term Cody-Waite reduction by pi/2 (exact for the small quotients that
// occur) followed by degree-7/8 minimax polynomials suffices: <= 2 ulp on |x| <= 64 (checked by
// tests/test_parity_gpu.py::test_sincos_accuracy), the same class as torch's CPU (Sleef u10) and CUDA
// kernels.  Arguments outside that range take the library sincosf.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sincos_heading(float x, float* sn, float* cs) {
  if (!(fabsf(x) <= 64.0f)) {  // also catches NaN/inf
    sincosf(x, sn, cs);
    return;
  }
  float q = rintf(x * 0.636619747f);  // round(x * 2/pi)
  int n = static_cast<int>(q);
  float r = fmaf(q, -1.57079601287841796875f, x);  // pi/2 split hi/mid/lo: products with small q are exact
  r = fmaf(q, -3.1391647326017846353352069854736328125e-7f, r);
  r = fmaf(q, -5.390302529957764765544681040410068817436695098876953125e-15f, r);
  float z = r * r;
  // sin(r), |r| <= pi/4
  float ps = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
  ps = fmaf(ps, z, -1.6666654611e-1f);
  float s = fmaf(ps * z, r, r);
  // cos(r), |r| <= pi/4
  float pc = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
  pc = fmaf(pc, z, 4.166664568298827e-2f);
  float c = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
  float s_out = (n & 1) ? c : s;
  float c_out = (n & 1) ? s : c;
  if (n & 2) s_out = -s_out;
  if ((n + 1) & 2) c_out = -c_out;
  *sn = s_out;
  *cs = c_out;
}

// ---------------------------------------------------------------------------------------------
// Angle wrap: (theta + pi) % (2 pi) - pi with torch.remainder semantics (fmod, then add the divisor
// when the result is non-zero and negative) -- robot_model.py:90.  For a = theta + pi in (-2pi, 4pi)
// fmod is a select between a, a - 2pi (exact, Sterbenz) and a itself, so the common case is branch-free.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float wrap_heading(float theta_raw) {
  float a = __fadd_rn(theta_raw, kPi);
  float m = a;
  if (a > -kTwoPi && a < 2.0f * kTwoPi) {
    m = (a >= kTwoPi) ? __fsub_rn(a, kTwoPi) : m;
    m = (a < 0.0f) ? __fadd_rn(a, kTwoPi) : m;
  } else {
    m = fmodf(a, kTwoPi);
    if (m != 0.0f && m < 0.0f) m = __fadd_rn(m, kTwoPi);
  }
  return __fsub_rn(m, kPi);
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// One unicycle step (UnicycleModel.transit, robot_model.py:75-95).  (x, y, th) is the clamped/wrapped
// state, tau its traversability; writes the raw successor (what the reference's in-place `+=` leaves in
// the input slot) and advances (x, y, th) to the clamped/wrapped successor.
struct Bounds {
  float u_min0, u_min1, u_max0, u_max1, dt;
};

__device__ __forceinline__ void unicycle_step(const GridGeom& g, const Bounds& b, float tau, float v0, float v1,
                                              float& x, float& y, float& th, float& xr, float& yr, float& thr) {
  v0 = clampf(v0, b.u_min0, b.u_max0);  // robot_model.py:82-83 (idempotent for already-clamped samples)
  v1 = clampf(v1, b.u_min1, b.u_max1);
  float sn, cs;
  sincos_heading(th, &sn, &cs);
  float tv = __fmul_rn(tau, v0);
  xr = __fadd_rn(x, __fmul_rn(__fmul_rn(tv, cs), b.dt));  // x += trav * v * cos(theta) * dt
  yr = __fadd_rn(y, __fmul_rn(__fmul_rn(tv, sn), b.dt));
  thr = __fadd_rn(th, __fmul_rn(__fmul_rn(tau, v1), b.dt));
  x = clampf(xr, g.x_min, g.x_max);
  y = clampf(yr, g.y_min, g.y_max);
  th = wrap_heading(thr);
}

// Stage/terminal cost term (objectives.py:46-53): ||p - goal|| + 1e4 * [tau <= thr].
__device__ __forceinline__ float goal_and_stuck_cost(float px, float py, float gx, float gy, float tau, float thr) {
  float dx = __fsub_rn(px, gx), dy = __fsub_rn(py, gy);
  float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  return __fadd_rn(d, (tau <= thr) ? kStuckPenalty : 0.0f);
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

// Two independent standard normals from two 32-bit words: u1 in (0,1], r = sqrt(-2 ln u1), angle = 2 pi u2.
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  float u1 = fmaf(static_cast<float>(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // a*2^-32 + 2^-33
  float u2 = fmaf(static_cast<float>(b), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

}  // namespace bnv
