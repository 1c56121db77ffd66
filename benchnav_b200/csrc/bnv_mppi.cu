// libbnvmppi.so -- host side of the C ABI declared in include/bnv_mppi.h.
// Owns the solver handle (device buffers, TMA descriptor, launch geometry) and launches the sm_100a
// kernels of mppi_kernels.cuh.  No PyTorch types anywhere: callers pass raw device pointers and a stream.
#include "bnv_internal.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <time.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "mppi_kernels.cuh"

namespace {

thread_local std::string g_last_error;

#define fail bnv_fail

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

constexpr size_t kMaxDynSmem = 220 * 1024;   // just below the 227 KB per-CTA limit
constexpr int kMaxPatchBytes = 64 * 1024;    // larger reach windows are looked up in the global map (L2)

// BNV_DEBUG_DISABLE bitmask (debugging aid): 1 = no TMA window (global-map lookups), 2 = no bulk noise load,
// 4 = no bulk recorded-state store.  All paths are regular product paths, selected otherwise by size/alignment.
unsigned debug_disable() {
  static const unsigned v = [] {
    const char* e = std::getenv("BNV_DEBUG_DISABLE");
    return e ? static_cast<unsigned>(std::strtoul(e, nullptr, 0)) : 0u;
  }();
  return v;
}

bool is_pow2_float(float v) {
  int e;
  return v > 0.0f && std::isfinite(v) && std::frexp(v, &e) == 0.5f;
}

}  // namespace

int bnv_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

struct bnv_mppi {
  bnv_mppi_cfg cfg{};
  int Kl = 0, k_offset = 0;
  int E = 1;           // environments per forward call (batch mode when > 1)
  bool stoch = false;  // stochastic-slip lookups
  bool problem_set = false, have_weights = false;
  uint64_t iteration = 0, launches = 0;
  bnv::EngineParams P{};
  // owned device buffers
  float* tau = nullptr;
  size_t tau_floats = 0;
  float* goals_dev = nullptr;  // [E][2]
  float* noise = nullptr;
  float* rec = nullptr;
  float* costs = nullptr;
  float* weights = nullptr;
  float* u_prev = nullptr;
  float* part_ms = nullptr;
  float* part_u = nullptr;
  float* shard_partial = nullptr;
  float* io_dev = nullptr;  // [3] state + [2T] u_out + [3(T+1)] opt states, staging for forward_host
  unsigned int* ticket = nullptr;
  unsigned int* ticket_grp = nullptr;    // [E][max_groups]
  float* part2_ms = nullptr;             // [E][max_groups][2]
  float* part2_u = nullptr;              // [E][max_groups][2T]
  int max_groups = 1;
  unsigned int* err_flag = nullptr;      // device word raised by a timed-out in-kernel wait
  float* stats = nullptr;
  float* replay = nullptr;             // [E][2T + 4] mean sequence + state of the last iteration (lean solvers' re-roll)
  const float* last_noise = nullptr;   // noise array of the last iteration: h->noise, or the caller's injected tensor
  unsigned int epoch = 0;
  int num_sms = 0;
  bool coop_ok = true;
  long long* dbg_ts = nullptr;
  // fused multi-GPU exchange: this rank's mailbox, the peers' mailboxes mapped through CUDA IPC
  float* mbox = nullptr;
  size_t mbox_floats = 0;
  std::vector<void*> peer_ptrs;       // host copy; entry [rank] is mbox itself
  float** peer_mbox_dev = nullptr;    // device array of the same pointers
  bool peers_attached = false;
  unsigned int xchg_seq = 0;
  float* merge_w = nullptr;  // dense copy of the candidate weights of bnv_mppi_merge_top
  size_t merge_w_cap = 0;
  int* top_idx = nullptr;
  unsigned long long* top_pairs = nullptr;
  size_t top_pairs_cap = 0, top_idx_cap = 0;
  unsigned int states_epoch = 0;       // launch whose optimal state sequence bnv_mppi_wait_states collects
  cudaStream_t states_stream = nullptr;
  float* io_host = nullptr;  // pinned mirror of io_dev
  float* io_host_dev = nullptr;  // its device-side address (zero-copy)
  unsigned long long* iter_dev = nullptr;  // device-resident iteration counter (graph-capturable launches)
  bool iter_external = false;              // ... advanced by the caller (bnv_closed_loop_step), no bump kernel
  // pre-launched iterations of forward_host (bnv_mppi_prelaunch)
  bool pre_enabled = false, pre_pending = false;
  cudaStream_t pre_stream = nullptr;        // internal stream of the pre-launched kernels
  unsigned int* pre_host = nullptr;         // pinned + mapped: [2 slots][4] {x, y, theta, seq}, [8 + slot] abort words
  unsigned int* pre_host_dev = nullptr;     // its device-side address
  unsigned int* pre_decision = nullptr;     // device word: the grid-wide go / abort decision of the pending launch
  unsigned int pre_seq = 0;                 // sequence number of the pending launch
  unsigned int pre_epoch = 0;               // its epoch (value of the completion / abort words)
  unsigned int pre_timeout_us = 2000;
  bool user_work = false;                   // work was queued on a caller's stream since the last pre-launched step
  cudaStream_t user_stream = nullptr;
  int grid = 0, warps = 0;
  bool wide = false;            // the wide (throughput) variant of the rollout kernel is selected (configure_launch)
  long long resident_ctas = 0;  // how many rollout CTAs the device can hold at once (cooperative-launch bound)
  bool fast_angles = false;
  size_t rollout_smem = 0, finalize_smem = 0;
  // optional CUDA-event timing of the rollout kernel alone (bench.py's roofline)
  bool timing = false;
  std::vector<cudaEvent_t> ev;  // pairs: [2i] before, [2i+1] after the rollout kernel
  size_t ev_used = 0;
};

namespace {

void free_all(bnv_mppi* h) {
  cudaFree(h->tau);
  cudaFree(h->goals_dev);
  cudaFree(h->iter_dev);
  cudaFree(h->pre_decision);
  if (h->pre_host) cudaFreeHost(h->pre_host);
  if (h->pre_stream) cudaStreamDestroy(h->pre_stream);
  cudaFree(h->noise);
  cudaFree(h->rec);
  cudaFree(h->costs);
  cudaFree(h->weights);
  cudaFree(h->u_prev);
  cudaFree(h->part_ms);
  cudaFree(h->part_u);
  cudaFree(h->shard_partial);
  cudaFree(h->io_dev);
  cudaFree(h->ticket);
  cudaFree(h->ticket_grp);
  cudaFree(h->part2_ms);
  cudaFree(h->part2_u);
  cudaFree(h->err_flag);
  cudaFree(h->stats);
  cudaFree(h->replay);
  cudaFree(h->dbg_ts);
  for (size_t r = 0; r < h->peer_ptrs.size(); ++r)
    if (h->peers_attached && static_cast<int>(r) != h->cfg.rank && h->peer_ptrs[r]) cudaIpcCloseMemHandle(h->peer_ptrs[r]);
  cudaFree(h->peer_mbox_dev);
  cudaFree(h->mbox);
  cudaFree(h->top_idx);
  cudaFree(h->merge_w);
  cudaFree(h->top_pairs);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  if (h->io_host) cudaFreeHost(h->io_host);
}

using RolloutFn = BnvRolloutFn;
using FinalizeFn = void (*)(bnv::EngineParams, const float*, int);

// rollout_kernel<kPatch, kPow2, kRecord, kFastAngles, kPhilox, false, false>: the single-solver instantiations
template <bool A, bool B, bool C, bool D>
RolloutFn pick_rollout4(bool e) {
  return e ? bnv::rollout_kernel<A, B, C, D, true, false, false> : bnv::rollout_kernel<A, B, C, D, false, false, false>;
}
template <bool A, bool B, bool C>
RolloutFn pick_rollout3(bool d, bool e) {
  return d ? pick_rollout4<A, B, C, true>(e) : pick_rollout4<A, B, C, false>(e);
}
template <bool A, bool B>
RolloutFn pick_rollout2(bool c, bool d, bool e) {
  return c ? pick_rollout3<A, B, true>(d, e) : pick_rollout3<A, B, false>(d, e);
}
template <bool A>
RolloutFn pick_rollout1(bool b, bool c, bool d, bool e) {
  return b ? pick_rollout2<A, true>(c, d, e) : pick_rollout2<A, false>(c, d, e);
}
RolloutFn pick_rollout(const bnv_mppi* h, bool philox) {
  const bool a = h->P.use_patch, b = h->P.geom.fast_grid, c = h->P.record, d = h->fast_angles;
  if (h->wide) {
    if (h->stoch || h->E > 1) return bnv_pick_rollout_wide_ext(a, b, philox, h->stoch, h->E > 1);
    return bnv_pick_rollout_wide(a, b, c, philox);
  }
  if (h->stoch || h->E > 1) return bnv_pick_rollout_ext(a, b, philox, h->stoch, h->E > 1);  // record + fast angles
  return a ? pick_rollout1<true>(b, c, d, philox) : pick_rollout1<false>(b, c, d, philox);
}
template <bool A, bool B>
FinalizeFn pick_finalize2(bool c) {
  return c ? bnv::finalize_kernel<A, B, true> : bnv::finalize_kernel<A, B, false>;
}
FinalizeFn pick_finalize(const bnv_mppi* h) {
  const bool a = h->P.use_patch, b = h->P.geom.fast_grid, c = h->fast_angles;
  if (a) return b ? pick_finalize2<true, true>(c) : pick_finalize2<true, false>(c);
  return b ? pick_finalize2<false, true>(c) : pick_finalize2<false, false>(c);
}

// Shared memory / occupancy of one candidate layout: sets the kernels' dynamic shared-memory attribute and returns
// how many CTAs the device holds at once (min over the Philox / injected-noise instantiations), 0 if it does not fit.
int layout_capacity(bnv_mppi* h, int w, int rec_split, size_t* smem_out, long long* cap_out) {
  bnv::EngineParams& P = h->P;
  bnv::RolloutSmem L = bnv::rollout_smem_layout(P.T, w, P.patch_w, P.patch_h, P.use_patch, P.record, h->stoch ? 2 : 1, rec_split,
                                                h->wide ? 1 : 0);
  *smem_out = L.total;
  *cap_out = 0;
  if (static_cast<size_t>(L.total) > kMaxDynSmem) return BNV_OK;
  long long cap = 0;
  for (bool philox : {false, true}) {
    const void* fn = reinterpret_cast<const void*>(pick_rollout(h, philox));
    // opt in to the full 220 KB once per kernel (a per-function limit, not a reservation): handles with different
    // horizons share the instantiations, so the limit must never shrink under a live handle
    BNV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaxDynSmem)));
    int per_sm = 0;
    BNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, w * 32, L.total));
    const long long c = static_cast<long long>(per_sm) * h->num_sms;
    cap = philox ? std::min(cap, c) : c;
  }
  *cap_out = cap;
  return BNV_OK;
}

// Choose warps per CTA so that noise + recorded-state slabs fit shared memory; prefer 4 (one per SM sub-partition),
// and ask the runtime how many CTAs can be resident at once (short horizons leave room for several CTAs per SM): a
// grid within that bound is launched cooperatively.
int configure_launch(bnv_mppi* h) {
  bnv::EngineParams& P = h->P;
  P.rec_split = 0;
  h->wide = false;
  // Latency variant first (4 warps per CTA, whole-horizon slabs, one CTA per SM at long horizons): it is the fastest
  // shape whenever its grid fits the device in one co-resident wave.  A single solver that does not fit runs the wide
  // variant (8 warps per CTA, two CTAs per SM, chunked staging): 4x the resident warps, bound by the FP32 pipe instead
  // of by the per-step dependency chain.  BNV_DEBUG_DISABLE bit 2048 = never wide, bit 4096 = always wide
  // (measurement aids; 4096 is also how the test suite drives every golden case through the wide variant).
  for (int w = bnv::kMaxWarps; w >= 1; w >>= 1) {
    size_t smem = 0;
    long long cap = 0;
    int rc = layout_capacity(h, w, 0, &smem, &cap);
    if (rc != BNV_OK) return rc;
    if (smem > kMaxDynSmem) continue;
    h->warps = w;
    h->grid = (h->Kl + w * 32 - 1) / (w * 32);
    const long long total = static_cast<long long>(h->grid) * h->E;
    // More CTAs than the device holds: flushing the recorded-state slab in two halves halves its footprint; adopt
    // that layout when the whole grid then fits the device at once (K = 32768 at T = 50: two CTAs per SM, one wave).
    const int tc = ((P.T + 1) / 2 + 1) & ~1;  // even split step, first half >= second half
    if (P.record && total > cap && tc >= 2 && tc <= P.T - 2 && !(debug_disable() & 1024u)) {
      size_t smem2 = 0;
      long long cap2 = 0;
      rc = layout_capacity(h, w, tc, &smem2, &cap2);
      if (rc != BNV_OK) return rc;
      if (cap2 >= total) {  // only when the launch becomes a single, co-resident wave (measured: 44.5 -> 37.8 us at
                            // K = 32768; with several waves either way the extra flush costs what the occupancy gains)
        P.rec_split = tc;
        smem = smem2;
        cap = cap2;
      }
    }
    h->rollout_smem = smem;
    h->resident_ctas = cap;
    // A single solver whose grid does not fit the device in one wave even so runs the wide variant (measured: K = 131072
    // at T = 50, 145 -> 100 us).  Batched solvers stay on the latency variant (64 environments x K = 4096: 88 vs 120 us --
    // their per-CTA fixed costs are paid once per wave, and the wide variant's waves are longer, not fewer).
    const bool want_wide = (debug_disable() & 4096u) || (h->E == 1 && total > cap);
    if (want_wide && h->fast_angles && !(debug_disable() & 2048u)) {
      h->wide = true;
      size_t smem_w = 0;
      long long cap_w = 0;
      rc = layout_capacity(h, bnv::kWideWarps, 0, &smem_w, &cap_w);
      if (rc != BNV_OK) return rc;
      if (smem_w <= kMaxDynSmem && cap_w > 0) {
        P.rec_split = 0;
        h->warps = bnv::kWideWarps;
        h->grid = (h->Kl + bnv::kWideWarps * 32 - 1) / (bnv::kWideWarps * 32);
        h->rollout_smem = smem_w;
        h->resident_ctas = 0;  // never launched cooperatively: the last CTA's epilogue serves every grid size
      } else {
        h->wide = false;  // (a reach window close to 64 KB leaves no room for two wide CTAs per SM)
      }
    }
    return BNV_OK;
  }
  return fail(BNV_ERR_UNSUPPORTED, "horizon %d does not fit the rollout kernel's shared-memory staging", P.T);
}

}  // namespace

struct PreStats {
  double t_sync = 0, t_post = 0, t_launch = 0, t_wait = 0;
  long calls = 0, posted = 0, fallbacks = 0, retries = 0;
};
static PreStats g_pre_stats;
static double now_us() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

extern "C" {
static int drain_prelaunch(bnv_mppi* h);
}
// Every entry point that touches the solver outside forward_host first cancels a pre-launched iteration that is still
// waiting for its state, and remembers that work now sits on the caller's stream.
#define BNV_DRAIN(h)                   \
  do {                                 \
    int rc__ = drain_prelaunch(h);     \
    if (rc__ != BNV_OK) return rc__;   \
  } while (0)
#define BNV_USER_WORK(h, s)                        \
  do {                                             \
    (h)->user_work = true;                         \
    (h)->user_stream = static_cast<cudaStream_t>(s); \
  } while (0)

extern "C" {

int bnv_abi_version(void) { return BNV_ABI_VERSION; }
const char* bnv_last_error(void) { return g_last_error.c_str(); }

int bnv_mppi_create(bnv_mppi** out, const bnv_mppi_cfg* cfg) {
  if (!out || !cfg) return fail(BNV_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->num_samples < 1 || cfg->horizon < 1) return fail(BNV_ERR_INVALID, "num_samples and horizon must be >= 1");
  if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size)
    return fail(BNV_ERR_INVALID, "rank %d / world_size %d invalid", cfg->rank, cfg->world_size);
  if (!(cfg->sigma[0] > 0.0f) || !(cfg->sigma[1] > 0.0f)) return fail(BNV_ERR_INVALID, "sigmas must be positive");
  if (!(cfg->lambda_ > 0.0f)) return fail(BNV_ERR_INVALID, "lambda_ must be positive");
  if (!(cfg->u_min[0] <= cfg->u_max[0]) || !(cfg->u_min[1] <= cfg->u_max[1]))
    return fail(BNV_ERR_INVALID, "u_min must not exceed u_max");
  if (!(cfg->dt > 0.0f)) return fail(BNV_ERR_INVALID, "dt must be positive");
  if (cfg->num_samples < cfg->world_size) return fail(BNV_ERR_INVALID, "fewer samples than shards");
  if (cfg->num_envs < 0 || cfg->num_envs > 65535) return fail(BNV_ERR_INVALID, "num_envs %d outside [0, 65535]", cfg->num_envs);
  const int E = cfg->num_envs > 1 ? cfg->num_envs : 1;
  const bool stoch = (cfg->flags & BNV_FLAG_STOCHASTIC_SLIP) != 0;
  if ((E > 1 || stoch) && cfg->world_size != 1)
    return fail(BNV_ERR_INVALID, "batched / stochastic-slip solvers need world_size == 1 (shard environments, not samples)");
  BNV_CUDA(cudaSetDevice(cfg->device));
  int major = 0;
  BNV_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cfg->device));
  if (major != 10) return fail(BNV_ERR_UNSUPPORTED, "device %d is sm_%dx; this library is built for sm_100a only", cfg->device, major);

  bnv_mppi* h = new (std::nothrow) bnv_mppi();
  if (!h) return fail(BNV_ERR_INVALID, "out of host memory");
  h->cfg = *cfg;
  h->E = E;
  h->stoch = stoch;
  const long long K = cfg->num_samples, W = cfg->world_size, r = cfg->rank;
  h->k_offset = static_cast<int>(r * K / W);  // shard = global samples [r K / W, (r+1) K / W)
  h->Kl = static_cast<int>((r + 1) * K / W) - h->k_offset;
  const int T = cfg->horizon, Kl = h->Kl;
  const bool record = (cfg->flags & BNV_FLAG_RECORD_STATES) != 0 || E > 1 || stoch;
  if ((E > 1 || stoch) && !(cfg->dt * std::max(std::fabs(cfg->u_min[1]), std::fabs(cfg->u_max[1])) < 3.0f)) {
    delete h;
    return fail(BNV_ERR_UNSUPPORTED, "batched / stochastic-slip solvers need dt * max|omega| < 3 rad per step");
  }
  // [E][3] states + [E][2T] u_out + [E][3(T+1)] opt states + completion words (padded to 4 floats)
  const size_t io_floats = static_cast<size_t>(E) * (3 + 2 * static_cast<size_t>(T) + 3 * static_cast<size_t>(T + 1)) + 4;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
  };
  const int max_grid = (Kl + 31) / 32;
  const size_t nE = static_cast<size_t>(E);  // every per-solver buffer carries a leading E
  alloc(reinterpret_cast<void**>(&h->noise), sizeof(float) * nE * Kl * T * 2);
  if (record) alloc(reinterpret_cast<void**>(&h->rec), sizeof(float) * nE * Kl * (T + 1) * 3);
  alloc(reinterpret_cast<void**>(&h->costs), sizeof(float) * nE * Kl);
  alloc(reinterpret_cast<void**>(&h->weights), sizeof(float) * nE * Kl);
  alloc(reinterpret_cast<void**>(&h->u_prev), sizeof(float) * nE * T * 2);
  alloc(reinterpret_cast<void**>(&h->part_ms), sizeof(float) * nE * max_grid * 2);
  alloc(reinterpret_cast<void**>(&h->part_u), sizeof(float) * nE * max_grid * 2 * T);
  alloc(reinterpret_cast<void**>(&h->shard_partial), sizeof(float) * (2 + 2 * T));
  alloc(reinterpret_cast<void**>(&h->io_dev), sizeof(float) * io_floats);
  alloc(reinterpret_cast<void**>(&h->ticket), 2 * nE * sizeof(unsigned int));
  h->max_groups = (max_grid + bnv::kMergeGroup - 1) / bnv::kMergeGroup;
  const size_t nG = static_cast<size_t>(h->max_groups);
  alloc(reinterpret_cast<void**>(&h->ticket_grp), nE * nG * sizeof(unsigned int));
  alloc(reinterpret_cast<void**>(&h->part2_ms), sizeof(float) * nE * nG * 2);
  alloc(reinterpret_cast<void**>(&h->part2_u), sizeof(float) * nE * nG * 2 * T);
  alloc(reinterpret_cast<void**>(&h->err_flag), sizeof(unsigned int));
  alloc(reinterpret_cast<void**>(&h->stats), 4 * nE * sizeof(float));
  if (!record) alloc(reinterpret_cast<void**>(&h->replay), nE * (2 * T + 4) * sizeof(float));
  alloc(reinterpret_cast<void**>(&h->goals_dev), 2 * nE * sizeof(float));
  if (cfg->world_size > 1) {  // mailbox: [2 parities][world ranks][2T columns] cells of three LL words {U[c] | M | S, tag}
    // + [2 parities] state cells of three LL words {x | y | theta, tag} (host-driven sharded solver)
    h->mbox_floats = 2 * static_cast<size_t>(cfg->world_size) * 2 * static_cast<size_t>(T) * 3 * 2 + 2 * 3 * 2;
    alloc(reinterpret_cast<void**>(&h->mbox), sizeof(float) * h->mbox_floats);
    alloc(reinterpret_cast<void**>(&h->peer_mbox_dev), sizeof(float*) * cfg->world_size);
    if (e == cudaSuccess) e = cudaMemset(h->mbox, 0, sizeof(float) * h->mbox_floats);
  }
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&h->io_host), sizeof(float) * io_floats, cudaHostAllocMapped);
  if (e == cudaSuccess) std::memset(h->io_host, 0, sizeof(float) * io_floats);
  if (e == cudaSuccess) e = cudaMemset(h->u_prev, 0, sizeof(float) * nE * T * 2);  // mppi.py:116
  if (e == cudaSuccess) e = cudaMemset(h->ticket, 0, 2 * nE * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(h->ticket_grp, 0, nE * nG * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(h->err_flag, 0, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(h->weights, 0, sizeof(float) * nE * Kl);    // mppi.py:126-128
  if (e == cudaSuccess && record) e = cudaMemset(h->rec, 0, sizeof(float) * nE * Kl * (T + 1) * 3);  // mppi.py:119-125
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    free_all(h);
    delete h;
    return fail(BNV_ERR_CUDA, "allocating solver buffers failed: %s", cudaGetErrorString(e));
  }
  bnv::EngineParams& P = h->P;
  P.bounds = {cfg->u_min[0], cfg->u_min[1], cfg->u_max[0], cfg->u_max[1], cfg->dt};
  P.lambda = cfg->lambda_;
  P.inv_lambda = 1.0f / cfg->lambda_;
  P.lambda_pow2 = is_pow2_float(cfg->lambda_) ? 1 : 0;
  // branch-free steps need the heading to move by less than pi per step (see mppi_math.cuh)
  h->fast_angles = cfg->dt * std::max(std::fabs(cfg->u_min[1]), std::fabs(cfg->u_max[1])) < 3.0f;
  P.icov0 = 1.0f / (cfg->sigma[0] * cfg->sigma[0]);  // inverse of diag(sigma^2), mppi.py:94-97
  P.icov1 = 1.0f / (cfg->sigma[1] * cfg->sigma[1]);
  P.sigma0 = cfg->sigma[0];
  P.sigma1 = cfg->sigma[1];
  P.seed_lo = static_cast<uint32_t>(cfg->seed);
  P.seed_hi = static_cast<uint32_t>(cfg->seed >> 32);
  P.k_offset = h->k_offset;
  P.rank = cfg->rank;
  P.peer_mbox = nullptr;
  P.num_envs = E;
  P.goals = nullptr;
  P.keep_mean = 1;
  P.done_flag = nullptr;
  P.iter_dev = nullptr;
  P.state_mailbox = nullptr;
  P.prelaunch_decision = nullptr;
  P.abort_flag = nullptr;
  P.xi_in = nullptr;
  P.xi_opt_in = nullptr;
  P.Kl = Kl;
  P.T = T;
  P.world = cfg->world_size;
  P.record = record ? 1 : 0;
  P.u_prev = h->u_prev;
  P.rec = h->rec;
  P.costs = h->costs;
  P.weights = h->weights;
  P.part_ms = h->part_ms;
  P.part_u = h->part_u;
  P.shard_partial = h->shard_partial;
  P.ticket = h->ticket;
  P.ticket_grp = h->ticket_grp;
  P.part2_ms = h->part2_ms;
  P.part2_u = h->part2_u;
  P.max_groups = h->max_groups;
  P.err_flag = h->err_flag;
  P.stats = h->stats;
  P.replay = h->replay;
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device);
  int coop_attr = 0;
  cudaDeviceGetAttribute(&coop_attr, cudaDevAttrCooperativeLaunch, cfg->device);
  h->coop_ok = coop_attr != 0 && !(debug_disable() & 64u);
  if (std::getenv("BNV_DEBUG_TS")) {
    if (cudaMalloc(reinterpret_cast<void**>(&h->dbg_ts), 24 * sizeof(long long)) == cudaSuccess)
      cudaMemset(h->dbg_ts, 0, 24 * sizeof(long long));
  }
  P.dbg_ts = h->dbg_ts;
  *out = h;
  return BNV_OK;
}

void bnv_mppi_destroy(bnv_mppi* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  drain_prelaunch(h);
  cudaDeviceSynchronize();
  free_all(h);
  delete h;
}

int bnv_mppi_set_problem(bnv_mppi* h, const float* risk_dev, int32_t grid_size, int32_t pitch, float resolution,
                         float x_min, float x_max, float y_min, float y_max, const float goal_xy[2],
                         float stuck_threshold, void* stream) {
  if (h && h->E > 1) return fail(BNV_ERR_INVALID, "batched solver: use bnv_mppi_set_problem_ex (one goal per environment)");
  if (h && h->stoch) return fail(BNV_ERR_INVALID, "stochastic-slip solver: use bnv_mppi_set_problem_ex (mean and std maps)");
  return bnv_mppi_set_problem_ex(h, risk_dev, nullptr, grid_size, pitch, 0, resolution, x_min, x_max, y_min, y_max,
                                 goal_xy, stuck_threshold, stream);
}

int bnv_mppi_set_problem_ex(bnv_mppi* h, const float* mean_dev, const float* std_dev, int32_t grid_size, int32_t pitch,
                            int64_t env_stride, float resolution, float x_min, float x_max, float y_min, float y_max,
                            const float* goals_xy, float stuck_threshold, void* stream) {
  if (!h || !mean_dev || !goals_xy) return fail(BNV_ERR_INVALID, "null argument");
  if (h->stoch != (std_dev != nullptr))
    return fail(BNV_ERR_INVALID, h->stoch ? "stochastic-slip solver needs a std map" : "std map given to a deterministic solver");
  if (grid_size < 1 || pitch < grid_size) return fail(BNV_ERR_INVALID, "grid_size %d / pitch %d invalid", grid_size, pitch);
  if (env_stride < 0) return fail(BNV_ERR_INVALID, "negative env_stride");
  if (!(resolution > 0.0f)) return fail(BNV_ERR_INVALID, "resolution must be positive");
  if (!(x_min < x_max) || !(y_min < y_max)) return fail(BNV_ERR_INVALID, "empty map limits");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  bnv::EngineParams& P = h->P;
  const int G = grid_size, E = h->E;
  const int cell = h->stoch ? 2 : 1;
  const int tpitch = (G + 3) & ~3;  // TMA needs a row stride that is a multiple of 16 bytes
  const size_t need = static_cast<size_t>(E) * G * tpitch * cell;
  BNV_CUDA(cudaStreamSynchronize(s));  // earlier iterations may still read the map / goals
  if (!h->tau || h->tau_floats != need) {
    if (h->tau) BNV_CUDA(cudaFree(h->tau));
    h->tau = nullptr;
    BNV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->tau), sizeof(float) * need));
    h->tau_floats = need;
  }
  dim3 grid((tpitch + 127) / 128, G, E);
  if (h->stoch)
    bnv::slip_map_kernel<<<grid, 128, 0, s>>>(mean_dev, std_dev, pitch, env_stride, reinterpret_cast<float2*>(h->tau), tpitch, G);
  else
    bnv::trav_map_kernel<<<grid, 128, 0, s>>>(mean_dev, pitch, env_stride, h->tau, tpitch, G);
  BNV_CUDA(cudaGetLastError());
  h->launches++;
  BNV_CUDA(cudaMemcpy(h->goals_dev, goals_xy, sizeof(float) * 2 * E, cudaMemcpyHostToDevice));

  P.tau = h->tau;
  P.G = G;
  P.pitch = tpitch;
  P.geom.x_min = x_min;
  P.geom.y_min = y_min;
  P.geom.x_max = x_max;
  P.geom.y_max = y_max;
  P.geom.res = resolution;
  P.geom.inv_res = 1.0f / resolution;
  P.geom.res_pow2 = is_pow2_float(resolution) ? 1 : 0;
  P.geom.fast_grid = (P.geom.res_pow2 && x_min == 0.0f && y_min == 0.0f) ? 1 : 0;
  P.goal_x = P.term_gx = goals_xy[0];
  P.goal_y = P.term_gy = goals_xy[1];
  P.goals = E > 1 ? h->goals_dev : nullptr;
  P.thr = stuck_threshold;

  // Reach bound: tau <= 1 and |v| <= vmax, so a rollout moves at most T * vmax * dt from the (clamped) start.
  const float vmax = std::max(std::fabs(h->cfg.u_min[0]), std::fabs(h->cfg.u_max[0]));
  const double reach_cells = static_cast<double>(P.T) * vmax * h->cfg.dt / resolution;
  const long long rho = static_cast<long long>(std::floor(reach_cells)) + 2;
  const long long side = 2 * rho + 1;
  const long long pw = (side + 3 + 3) & ~3LL;  // +3: the window's x origin is rounded down to a multiple of 4 cells
  P.use_patch = (pw * cell <= 256 && side <= 256 && pw * side * 4 * cell <= kMaxPatchBytes) ? 1 : 0;
  if (debug_disable() & 1u) P.use_patch = 0;
  if (P.use_patch) {
    P.rho = static_cast<int>(rho);
    P.patch_w = static_cast<int>(pw);
    P.patch_h = static_cast<int>(side);
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(BNV_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    // the E maps are stacked along y (environment e = rows [e G, (e+1) G)); a stochastic cell is two floats wide
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(G) * cell, static_cast<cuuint64_t>(G) * E};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(tpitch) * cell * sizeof(float)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(P.patch_w * cell), static_cast<cuuint32_t>(P.patch_h)};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = enc(&P.tau_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h->tau, gdim, gstride, box, estride,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(BNV_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  } else {
    P.rho = 0;
    P.patch_w = P.patch_h = 0;
    std::memset(&P.tau_map, 0, sizeof(P.tau_map));
  }
  int rc = configure_launch(h);
  if (rc != BNV_OK) return rc;
  // the launch geometry may have changed: restart the arrival counters (the stream was synchronised above)
  BNV_CUDA(cudaMemsetAsync(h->ticket, 0, 2 * static_cast<size_t>(E) * sizeof(unsigned int), s));
  BNV_CUDA(cudaMemsetAsync(h->ticket_grp, 0, static_cast<size_t>(E) * h->max_groups * sizeof(unsigned int), s));
  h->finalize_smem = 128 + (P.use_patch ? ((P.patch_w * P.patch_h + 31) / 32) * 32 * 4 : 0) + 2 * (2 * P.T + 4) * 4 + 16;
  if (!h->stoch && E == 1)
    BNV_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_finalize(h)),
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaxDynSmem)));
  h->problem_set = true;
  return BNV_OK;
}

static int launch_forward(bnv_mppi* h, const float* state_dev, const float* state_host, const float* noise_dev,
                          float* u_out_dev, float* opt_states_dev, cudaStream_t s, const float* xi_dev = nullptr,
                          const float* xi_opt_dev = nullptr, unsigned int mailbox_seq = 0, int state_role = 0) {
  if (state_host && h->E > 1) return fail(BNV_ERR_INVALID, "batched solver: states must be device-resident [E,3]");
  if (state_host) {  // state travels by value in the launch packet; remembered for finalize (world_size > 1)
    for (int i = 0; i < 3; ++i) h->P.state_val[i] = state_host[i];
    h->P.state_inline = 1;
  } else {
    h->P.state_inline = 0;
  }
  bnv::EngineParams P = h->P;
  P.state_role = state_role;
  P.dbg_flags = debug_disable() >> 14;
  const bool philox = noise_dev == nullptr;
  P.noise_in = noise_dev;
  P.noise_out = h->noise;
  h->last_noise = philox ? h->noise : noise_dev;
  P.xi_in = xi_dev;
  P.xi_opt_in = xi_opt_dev;
  P.iter_lo = static_cast<uint32_t>(h->iteration);
  P.iter_hi = static_cast<uint32_t>(h->iteration >> 32);
  P.noise_bulk_ok = (!philox && (reinterpret_cast<uintptr_t>(noise_dev) & 15u) == 0) ? 1 : 0;
  P.rec_bulk_ok = (reinterpret_cast<uintptr_t>(P.rec) & 15u) == 0 ? 1 : 0;
  if (debug_disable() & 2u) P.noise_bulk_ok = 0;
  if (debug_disable() & 4u) P.rec_bulk_ok = 0;
  P.state = state_dev;
  h->P.state = state_dev;  // finalize (world_size > 1) re-reads the state of the iteration in flight
  P.u_out = u_out_dev;
  P.opt_rec = opt_states_dev;
  if (mailbox_seq != 0u) {  // pre-launched: the state arrives later through the host-mapped mailbox
    P.state_mailbox = h->pre_host_dev;
    P.mailbox_seq = mailbox_seq;
    P.mailbox_timeout_us = h->pre_timeout_us;
    P.prelaunch_decision = h->pre_decision;
    P.abort_flag = h->pre_host_dev + 8 + (mailbox_seq & 1u);  // one abort word per mailbox slot
    BNV_CUDA(cudaMemsetAsync(h->pre_decision, 0, sizeof(unsigned int), s));
  }
  // The grid is co-resident iff it has at most resident_ctas CTAs (occupancy x SMs; one CTA per SM at long
  // horizons).  Then launch cooperatively (residency guaranteed by the driver) and use the deferred-store epilogue.
  const bool coop = h->coop_ok && static_cast<long long>(h->grid) * h->E <= h->resident_ctas;
  h->epoch = (h->epoch == 0xFFFFFFFFu) ? 1u : h->epoch + 1u;
  P.epoch = h->epoch;
  P.coop = coop ? 1 : 0;
  if (h->peers_attached) {
    h->xchg_seq = (h->xchg_seq == 0xFFFFFFFFu) ? 1u : h->xchg_seq + 1u;  // advances in lock-step on every rank
    P.xchg_seq = h->xchg_seq;
    P.peer_mbox = h->peer_mbox_dev;
    for (size_t r = 0; r < 8; ++r) P.peer_mbox_val[r] = r < h->peer_ptrs.size() ? static_cast<float*>(h->peer_ptrs[r]) : nullptr;
  }
  const bool timed = h->timing && h->ev_used + 2 <= h->ev.size();
  if (timed) BNV_CUDA(cudaEventRecord(h->ev[h->ev_used], s));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(h->grid, h->E);
  cfg.blockDim = dim3(h->warps * 32);
  cfg.dynamicSmemBytes = h->rollout_smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (coop && !(debug_disable() & 128u)) ? 1 : 0;  // bit 128: measurement aid, plain launch of the coop path
  BNV_CUDA(cudaLaunchKernelEx(&cfg, pick_rollout(h, philox), P));
  h->launches++;
  if (!coop) {  // the grid was not co-resident: its weights are normalised by a second, fully parallel kernel
    int spb_shift = 0;
    while ((1 << spb_shift) < h->warps * 32) ++spb_shift;
    bnv::normalize_weights_kernel<<<dim3((h->Kl + 255) / 256, h->E), 256, 0, s>>>(
        h->weights, h->part_ms, h->stats, h->Kl, spb_shift, h->grid, mailbox_seq != 0u ? h->pre_decision : nullptr);
    BNV_CUDA(cudaGetLastError());
  }
  if (timed) {
    BNV_CUDA(cudaEventRecord(h->ev[h->ev_used + 1], s));
    h->ev_used += 2;
  }
  if (!coop) h->launches++;
  if (h->iter_dev && h->P.iter_dev) {  // the counter advances on the device, in stream order (and on every graph replay)
    if (!h->iter_external) {
      bnv::bump_iteration_kernel<<<1, 1, 0, s>>>(h->iter_dev);
      BNV_CUDA(cudaGetLastError());
      h->launches++;
    }
  } else if (philox) {
    h->iteration++;
  }
  h->have_weights = (P.world == 1) || h->peers_attached;
  return BNV_OK;
}

// Cancel a pre-launched iteration that is still waiting for its state (every entry point that touches the solver
// outside forward_host calls this first): mark its mailbox slot, wait for the launch to abort, and give back the
// iteration number it had taken, so that the noise stream is the same as without pre-launching.
static int drain_prelaunch(bnv_mppi* h) {
  if (!h->pre_pending) return BNV_OK;
  volatile unsigned int* slot = h->pre_host + 4u * (h->pre_seq & 1u);
  slot[3] = 0xFFFFFFFFu;
  std::atomic_thread_fence(std::memory_order_seq_cst);
  BNV_CUDA(cudaStreamSynchronize(h->pre_stream));
  h->pre_pending = false;
  if (h->iteration > 0) h->iteration--;
  return BNV_OK;
}
int bnv_mppi_prelaunch(bnv_mppi* h, int32_t enable, uint32_t timeout_us) {
  if (!h) return fail(BNV_ERR_INVALID, "null argument");
  if (h->cfg.world_size != 1 || h->E != 1) return fail(BNV_ERR_INVALID, "pre-launching needs a single, unsharded solver");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  if (enable) {
    if (!h->pre_stream) BNV_CUDA(cudaStreamCreateWithFlags(&h->pre_stream, cudaStreamNonBlocking));
    if (!h->pre_host) {
      BNV_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->pre_host), 16 * sizeof(unsigned int), cudaHostAllocMapped));
      std::memset(h->pre_host, 0, 16 * sizeof(unsigned int));
      BNV_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->pre_host_dev), h->pre_host, 0));
    }
    if (!h->pre_decision) BNV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->pre_decision), 8 * sizeof(unsigned int)));
    h->pre_timeout_us = timeout_us ? timeout_us : 2000u;
  } else if (h->pre_stream) {
    BNV_CUDA(cudaStreamSynchronize(h->pre_stream));
  }
  h->pre_enabled = enable != 0;
  if (!enable && std::getenv("BNV_DEBUG_PRE") && g_pre_stats.calls > 0) {
    const PreStats& p = g_pre_stats;
    std::fprintf(stderr, "[bnv prelaunch] calls %ld posted %ld fallbacks %ld retries %ld | per call us: sync %.2f post/plain %.2f "
                 "launch-next %.2f wait %.2f\n", p.calls, p.posted, p.fallbacks, p.retries, p.t_sync / p.calls,
                 p.t_post / p.calls, p.t_launch / p.calls, p.t_wait / p.calls);
    g_pre_stats = PreStats();
  }
  return BNV_OK;
}

int bnv_mppi_forward(bnv_mppi* h, const float* state_dev, const float* noise_dev, float* u_out_dev,
                     float* opt_states_dev, void* stream) {
  if (!h || !state_dev) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->problem_set) return fail(BNV_ERR_STATE, "bnv_mppi_set_problem must be called before forward");
  if ((h->cfg.world_size == 1 || h->peers_attached) && (!u_out_dev || !opt_states_dev))
    return fail(BNV_ERR_INVALID, "null output buffer");
  if (h->stoch && noise_dev) return fail(BNV_ERR_INVALID, "stochastic-slip solver: inject noise through bnv_mppi_forward_ex");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  return launch_forward(h, state_dev, nullptr, noise_dev, u_out_dev, opt_states_dev, static_cast<cudaStream_t>(stream));
}

int bnv_mppi_forward_ex(bnv_mppi* h, const float* state_dev, const float* noise_dev, const float* xi_dev,
                        const float* xi_opt_dev, float* u_out_dev, float* opt_states_dev, void* stream) {
  if (!h || !state_dev || !u_out_dev || !opt_states_dev) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->problem_set) return fail(BNV_ERR_STATE, "bnv_mppi_set_problem must be called before forward");
  if (h->stoch) {
    const bool all = noise_dev && xi_dev && xi_opt_dev, none = !noise_dev && !xi_dev && !xi_opt_dev;
    if (!all && !none) return fail(BNV_ERR_INVALID, "noise_dev, xi_dev and xi_opt_dev must be all given or all NULL");
  } else if (xi_dev || xi_opt_dev) {
    return fail(BNV_ERR_INVALID, "lookup normals given to a deterministic solver");
  }
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  return launch_forward(h, state_dev, nullptr, noise_dev, u_out_dev, opt_states_dev, static_cast<cudaStream_t>(stream),
                        xi_dev, xi_opt_dev);
}

int bnv_mppi_forward_follow(bnv_mppi* h, const float* noise_dev, float* u_out_dev, float* opt_states_dev, void* stream) {
  if (!h || !u_out_dev || !opt_states_dev) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->problem_set) return fail(BNV_ERR_STATE, "bnv_mppi_set_problem must be called before forward");
  if (h->cfg.world_size < 2 || !h->peers_attached)
    return fail(BNV_ERR_INVALID, "forward_follow needs a sharded solver with attached peers");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  h->P.state = nullptr;
  return launch_forward(h, nullptr, nullptr, noise_dev, u_out_dev, opt_states_dev, static_cast<cudaStream_t>(stream),
                        nullptr, nullptr, 0u, 2);
}

int bnv_mppi_forward_state(bnv_mppi* h, const float state_host[3], const float* noise_dev, float* u_out_dev,
                           float* opt_states_dev, void* stream) {
  if (!h || !state_host) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->problem_set) return fail(BNV_ERR_STATE, "bnv_mppi_set_problem must be called before forward");
  if ((h->cfg.world_size == 1 || h->peers_attached) && (!u_out_dev || !opt_states_dev))
    return fail(BNV_ERR_INVALID, "null output buffer");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  return launch_forward(h, nullptr, state_host, noise_dev, u_out_dev, opt_states_dev, static_cast<cudaStream_t>(stream));
}

// Spin on the completion word of launch `epoch` (and, for a pre-launched one, on its abort word).  Returns 1 when the
// results are in the staging buffer, 0 when the launch aborted, negative on error.
static int wait_host_results(bnv_mppi* h, cudaStream_t s, unsigned int epoch, bool may_abort, unsigned int seq = 0,
                             int stage = 0) {  // stage 1: only u* is awaited (the word raised by signal_action)
  const int T = h->P.T;
  const size_t flag_off = 3 + 2 * static_cast<size_t>(T) + 3 * static_cast<size_t>(T + 1) + (stage ? 1 : 0);
  volatile unsigned int* flag = reinterpret_cast<volatile unsigned int*>(h->io_host + flag_off);
  volatile unsigned int* aborted = h->pre_host ? h->pre_host + 8 + (seq & 1u) : nullptr;  // the awaited launch's own word
  bool seen = false;
  for (long spins = 0; spins < 50000000L; ++spins) {  // a fault or a stuck device ends in the synchronise below
    if (*flag == epoch) {
      seen = true;
      break;
    }
    if (may_abort && *aborted == epoch) return 0;
    // a rare liveness check (about once a millisecond): a stream query costs a microsecond or two, which must not
    // land inside the ~20 us the kernel normally takes
    if ((spins & 0xFFFF) == 0xFFFF && !may_abort && cudaStreamQuery(s) != cudaErrorNotReady) break;
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  if (!seen) {
    BNV_CUDA(cudaStreamSynchronize(s));
    if (*flag != epoch) return (may_abort && *aborted == epoch) ? 0 : fail(BNV_ERR_CUDA, "the iteration did not complete");
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return 1;
}

// forward_host with pre-launching (bnv_mppi_prelaunch): every call posts its state to the kernel that the PREVIOUS call
// queued -- already resident, its prologue done, polling the host-mapped mailbox -- and queues the next iteration's
// kernel behind it before waiting, so that neither the launch call nor the launch latency is on the step's critical
// path.  A launch whose state does not arrive within the timeout aborts itself (every CTA follows one grid-wide
// decision) and the step falls back to a plain launch; iteration numbers are given back, so the noise stream is the
// one of the plain path.
static int forward_host_prelaunched(bnv_mppi* h, const float state_host[3], float* u_out_host, float* opt_states_host,
                                    int depth) {
  const int T = h->P.T;
  const double t_a = now_us();
  g_pre_stats.calls++;
  cudaStream_t ps = h->pre_stream;
  if (h->user_work) {  // order the internal stream after whatever the caller queued on its own stream
    BNV_CUDA(cudaStreamSynchronize(h->user_stream));
    h->user_work = false;
  }
  if (!h->io_host_dev) BNV_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->io_host_dev), h->io_host, 0));
  float* out_dev = h->io_host_dev;
  const size_t flag_off = 3 + 2 * static_cast<size_t>(T) + 3 * static_cast<size_t>(T + 1);
  volatile unsigned int* aborted = h->pre_host + 8 + (h->pre_seq & 1u);  // abort word of the pending launch's slot
  unsigned int cur_epoch = 0;
  const unsigned int cur_seq = h->pre_seq;
  bool posted = false;
  if (h->pre_pending && *aborted == h->pre_epoch) {  // the waiting launch timed out before this call
    BNV_CUDA(cudaStreamSynchronize(ps));
    h->pre_pending = false;
    if (h->iteration > 0) h->iteration--;
    g_pre_stats.fallbacks++;
  }
  const double t_b = now_us();
  h->P.done_flag = reinterpret_cast<unsigned int*>(out_dev + flag_off);
  if (h->pre_pending) {
    volatile unsigned int* slot = h->pre_host + 4u * (h->pre_seq & 1u);
    unsigned int bits[3];
    std::memcpy(bits, state_host, sizeof(bits));
    slot[0] = bits[0];
    slot[1] = bits[1];
    slot[2] = bits[2];
    std::atomic_thread_fence(std::memory_order_release);
    slot[3] = h->pre_seq;
    cur_epoch = h->pre_epoch;
    posted = true;
    g_pre_stats.posted++;
  } else {
    int rc = launch_forward(h, nullptr, state_host, nullptr, out_dev + 3, out_dev + 3 + 2 * T, ps);
    if (rc != BNV_OK) {
      h->P.done_flag = nullptr;
      return rc;
    }
    cur_epoch = h->epoch;
  }
  const double t_c = now_us();
  // queue the next iteration behind the current one
  unsigned int next_seq = h->pre_seq + 1u;
  if (next_seq == 0u || next_seq == 0xFFFFFFFFu) next_seq = 1u;
  h->pre_host[4u * (next_seq & 1u) + 3u] = 0u;  // the slot last carried the sequence number before the previous one
  std::atomic_thread_fence(std::memory_order_seq_cst);
  int rc = launch_forward(h, nullptr, nullptr, nullptr, out_dev + 3, out_dev + 3 + 2 * T, ps, nullptr, nullptr, next_seq);
  h->P.done_flag = nullptr;
  if (rc != BNV_OK) return rc;
  h->pre_seq = next_seq;
  h->pre_epoch = h->epoch;
  h->pre_pending = true;
  const double t_d = now_us();
  const int got = wait_host_results(h, ps, cur_epoch, posted, cur_seq, opt_states_host ? 0 : 1);
  const double t_e = now_us();
  g_pre_stats.t_sync += t_b - t_a;
  g_pre_stats.t_post += t_c - t_b;
  g_pre_stats.t_launch += t_d - t_c;
  g_pre_stats.t_wait += t_e - t_d;
  if (got < 0) return got;
  if (got == 0) {
    // the posted launch aborted in the same instant: cancel the queued one, give both iteration numbers back, retry
    g_pre_stats.retries++;
    if (depth > 2) return fail(BNV_ERR_CUDA, "pre-launched iterations keep aborting");
    rc = drain_prelaunch(h);
    if (rc != BNV_OK) return rc;
    if (h->iteration > 0) h->iteration--;
    return forward_host_prelaunched(h, state_host, u_out_host, opt_states_host, depth + 1);
  }
  std::memcpy(u_out_host, h->io_host + 3, 2 * static_cast<size_t>(T) * sizeof(float));
  h->states_epoch = cur_epoch;
  h->states_stream = ps;
  if (opt_states_host) std::memcpy(opt_states_host, h->io_host + 3 + 2 * T, 3 * static_cast<size_t>(T + 1) * sizeof(float));
  return BNV_OK;
}

static int forward_host_impl(bnv_mppi* h, const float state_host[3], const float* noise_dev, float* u_out_host,
                             float* opt_states_host, void* stream);

int bnv_mppi_forward_host(bnv_mppi* h, const float state_host[3], const float* noise_dev, float* u_out_host,
                          float* opt_states_host, void* stream) {
  if (!opt_states_host) return fail(BNV_ERR_INVALID, "null argument");
  return forward_host_impl(h, state_host, noise_dev, u_out_host, opt_states_host, stream);
}

int bnv_mppi_forward_host_action(bnv_mppi* h, const float state_host[3], float* u_out_host, void* stream) {
  return forward_host_impl(h, state_host, nullptr, u_out_host, nullptr, stream);
}

int bnv_mppi_wait_states(bnv_mppi* h, float* opt_states_host) {
  if (!h || !opt_states_host) return fail(BNV_ERR_INVALID, "null argument");
  if (h->states_epoch == 0u) return fail(BNV_ERR_STATE, "wait_states needs a preceding forward_host / forward_host_action");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  const int got = wait_host_results(h, h->states_stream, h->states_epoch, false, 0u, 0);
  if (got < 0) return got;
  const int T = h->P.T;
  std::memcpy(opt_states_host, h->io_host + 3 + 2 * T, 3 * static_cast<size_t>(T + 1) * sizeof(float));
  return BNV_OK;
}

// opt_states_host == nullptr: return as soon as u* is in host memory (bnv_mppi_forward_host_action)
static int forward_host_impl(bnv_mppi* h, const float state_host[3], const float* noise_dev, float* u_out_host,
                             float* opt_states_host, void* stream) {
  if (!h || !state_host || !u_out_host) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->problem_set) return fail(BNV_ERR_STATE, "bnv_mppi_set_problem must be called before forward");
  if (h->E != 1) return fail(BNV_ERR_INVALID, "forward_host needs a single environment");
  if (h->cfg.world_size != 1 && !h->peers_attached)
    return fail(BNV_ERR_INVALID, "forward_host on a sharded solver needs attached peers (bnv_mppi_attach_peers); the "
                                 "other ranks call bnv_mppi_forward_follow");
  if (h->stoch && noise_dev) return fail(BNV_ERR_INVALID, "stochastic-slip solver: inject noise through bnv_mppi_forward_ex");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  if (h->pre_enabled && !noise_dev && !h->P.iter_dev)
    return forward_host_prelaunched(h, state_host, u_out_host, opt_states_host, 0);
  BNV_DRAIN(h);
  const int T = h->P.T;
  // Host -> device: the 12-byte state rides in the kernel's launch packet.  Device -> host: the kernel stores u* and
  // the optimal state sequence straight into the handle's pinned, device-mapped staging buffer (zero-copy over
  // PCIe) and raises a completion word as soon as both are written (before its tail: slab stores draining, CTAs
  // exiting); the host polls that word instead of paying a stream synchronisation.
  if (!h->io_host_dev) BNV_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->io_host_dev), h->io_host, 0));
  float* out_dev = h->io_host_dev;
  const size_t flag_off = 3 + 2 * static_cast<size_t>(T) + 3 * static_cast<size_t>(T + 1);
  h->P.done_flag = reinterpret_cast<unsigned int*>(out_dev + flag_off);
  int rc = launch_forward(h, nullptr, state_host, noise_dev, out_dev + 3, out_dev + 3 + 2 * T, s, nullptr, nullptr, 0u,
                          h->cfg.world_size > 1 ? 1 : 0);  // sharded: this rank leads, its kernel broadcasts the state
  h->P.done_flag = nullptr;
  if (rc != BNV_OK) return rc;
  h->user_work = true;
  h->user_stream = s;
  const int got = wait_host_results(h, s, h->epoch, false, 0u, opt_states_host ? 0 : 1);
  if (got < 0) return got;
  std::memcpy(u_out_host, h->io_host + 3, 2 * static_cast<size_t>(T) * sizeof(float));
  h->states_epoch = h->epoch;
  h->states_stream = s;
  if (opt_states_host) std::memcpy(opt_states_host, h->io_host + 3 + 2 * T, 3 * static_cast<size_t>(T + 1) * sizeof(float));
  return BNV_OK;
}

int bnv_mppi_forward_host_batch(bnv_mppi* h, const float* states_host, float* u_out_host, float* opt_states_host,
                                void* stream) {
  if (!h || !states_host || !u_out_host || !opt_states_host) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->problem_set) return fail(BNV_ERR_STATE, "bnv_mppi_set_problem must be called before forward");
  if (h->cfg.world_size != 1) return fail(BNV_ERR_INVALID, "forward_host_batch needs an unsharded solver");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  // One staged copy of the [E][3] states up, one synchronisation.  (The kernels read the states by many CTAs per environment: they belong in HBM, not behind PCIe.)
  const size_t nE = static_cast<size_t>(h->E), T = static_cast<size_t>(h->P.T);
  const size_t n_st = 3 * nE, n_u = 2 * T * nE, n_o = 3 * (T + 1) * nE;
  std::memcpy(h->io_host, states_host, n_st * sizeof(float));
  BNV_CUDA(cudaMemcpyAsync(h->io_dev, h->io_host, n_st * sizeof(float), cudaMemcpyHostToDevice, s));
  if (!h->io_host_dev) BNV_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->io_host_dev), h->io_host, 0));
  // results: stored by each environment's last CTA straight into the pinned, device-mapped staging buffer (zero-copy
  // over PCIe, as in bnv_mppi_forward_host) -- no copy behind the kernel
  const int rc = launch_forward(h, h->io_dev, nullptr, nullptr, h->io_host_dev + n_st, h->io_host_dev + n_st + n_u, s);
  if (rc != BNV_OK) return rc;
  BNV_CUDA(cudaStreamSynchronize(s));
  std::memcpy(u_out_host, h->io_host + n_st, n_u * sizeof(float));
  std::memcpy(opt_states_host, h->io_host + n_st + n_u, n_o * sizeof(float));
  return BNV_OK;
}

int bnv_mppi_mailbox_handle(bnv_mppi* h, unsigned char out[64]) {
  if (!h || !out) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->mbox) return fail(BNV_ERR_STATE, "no mailbox: world_size is 1");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  cudaIpcMemHandle_t hd;
  BNV_CUDA(cudaIpcGetMemHandle(&hd, h->mbox));
  std::memcpy(out, &hd, 64);
  return BNV_OK;
}

int bnv_mppi_attach_peers(bnv_mppi* h, const unsigned char* handles) {
  if (!h || !handles) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->mbox) return fail(BNV_ERR_STATE, "no mailbox: world_size is 1");
  if (h->peers_attached) return fail(BNV_ERR_STATE, "peers already attached");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  const int W = h->cfg.world_size;
  h->peer_ptrs.assign(W, nullptr);
  for (int r = 0; r < W; ++r) {
    if (r == h->cfg.rank) {
      h->peer_ptrs[r] = h->mbox;
      continue;
    }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handles + 64 * static_cast<size_t>(r), 64);
    cudaError_t e = cudaIpcOpenMemHandle(&h->peer_ptrs[r], hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < r; ++q)
        if (q != h->cfg.rank && h->peer_ptrs[q]) cudaIpcCloseMemHandle(h->peer_ptrs[q]);
      h->peer_ptrs.clear();
      return fail(BNV_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
    }
  }
  BNV_CUDA(cudaMemcpy(h->peer_mbox_dev, h->peer_ptrs.data(), sizeof(float*) * W, cudaMemcpyHostToDevice));
  h->peers_attached = true;
  h->xchg_seq = 0;
  return BNV_OK;
}

const float* bnv_mppi_partial(const bnv_mppi* h) { return h ? h->shard_partial : nullptr; }
int32_t bnv_mppi_partial_len(const bnv_mppi* h) { return h ? 2 + 2 * h->P.T : 0; }

int bnv_mppi_finalize(bnv_mppi* h, const float* gathered_partials_dev, float* u_out_dev, float* opt_states_dev,
                      void* stream) {
  if (!h || !gathered_partials_dev || !u_out_dev || !opt_states_dev) return fail(BNV_ERR_INVALID, "null argument");
  if (h->peers_attached) return fail(BNV_ERR_STATE, "peers attached: forward already exchanged and finalised");
  if (!h->problem_set) return fail(BNV_ERR_STATE, "finalize needs a preceding forward");
  if (!h->P.state && !h->P.state_inline) return fail(BNV_ERR_STATE, "finalize needs a preceding forward");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  bnv::EngineParams P = h->P;
  P.u_out = u_out_dev;
  P.opt_rec = opt_states_dev;
  pick_finalize(h)<<<1, bnv::kFinalizeThreads, h->finalize_smem, static_cast<cudaStream_t>(stream)>>>(
      P, gathered_partials_dev, h->cfg.rank);
  BNV_CUDA(cudaGetLastError());
  h->launches++;
  h->have_weights = true;
  return BNV_OK;
}

int bnv_mppi_top_samples(bnv_mppi* h, int32_t n, float* states_out_dev, float* weights_out_dev, void* stream) {
  if (!h || !states_out_dev || !weights_out_dev) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->have_weights) return fail(BNV_ERR_STATE, "top_samples needs a completed forward");
  if (!h->P.record && (!h->replay || !h->last_noise))
    return fail(BNV_ERR_STATE, "top_samples on a solver without recorded states needs a completed forward");
  if (n < 1 || n > h->Kl) return fail(BNV_ERR_INVALID, "num_samples %d outside [1, %d]", n, h->Kl);  // mppi.py:229
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  int n_pad = 1;
  while (n_pad < n) n_pad <<= 1;
  const size_t nE = static_cast<size_t>(h->E);
  if (h->top_idx_cap < nE * n) {
    BNV_CUDA(cudaStreamSynchronize(s));
    if (h->top_idx) BNV_CUDA(cudaFree(h->top_idx));
    h->top_idx = nullptr;
    BNV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->top_idx), sizeof(int) * nE * n));
    h->top_idx_cap = nE * n;
  }
  unsigned long long* pairs_global = nullptr;
  size_t smem = static_cast<size_t>(n_pad) * 8;
  if (n_pad > bnv::kTopnSmemPairs) {
    if (h->top_pairs_cap < nE * n_pad) {
      BNV_CUDA(cudaStreamSynchronize(s));
      if (h->top_pairs) BNV_CUDA(cudaFree(h->top_pairs));
      h->top_pairs = nullptr;
      BNV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->top_pairs), sizeof(unsigned long long) * nE * n_pad));
      h->top_pairs_cap = nE * n_pad;
    }
    pairs_global = h->top_pairs;
    smem = 0;
  }
  BNV_CUDA(cudaFuncSetAttribute(bnv::topn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                bnv::kTopnSmemPairs * 8));
  const bool fused_gather = h->P.record && n <= bnv::kTopnFusedGatherMax;
  bnv::topn_select_kernel<<<h->E, bnv::kTopnThreads, smem, s>>>(h->weights, h->Kl, n, n_pad, pairs_global,
                                                                 weights_out_dev, h->top_idx,
                                                                 fused_gather ? h->rec : nullptr, 3 * (h->P.T + 1),
                                                                 states_out_dev);
  BNV_CUDA(cudaGetLastError());
  if (fused_gather) {
    h->launches += 1;
    return BNV_OK;
  }
  if (h->P.record) {
    bnv::gather_rows_kernel<<<dim3(n, h->E), 128, 0, s>>>(h->rec, h->top_idx, 3 * (h->P.T + 1), h->Kl, states_out_dev);
  } else {
    // lean solver (no recorded states): roll the n selected samples out again from the iteration's saved start
    // (E == 1: batched and stochastic solvers always record)
    bnv::EngineParams P = h->P;
    const bool pow2 = P.geom.fast_grid != 0;
    using RerollFn = void (*)(bnv::EngineParams, const float*, const float*, const int*, int, float*);
    RerollFn fn = pow2 ? (h->fast_angles ? bnv::reroll_kernel<true, true> : bnv::reroll_kernel<true, false>)
                       : (h->fast_angles ? bnv::reroll_kernel<false, true> : bnv::reroll_kernel<false, false>);
    const size_t smem_r = static_cast<size_t>(P.T) * 16;
    if (smem_r > 48 * 1024) BNV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_r)));
    fn<<<(n + 127) / 128, 128, smem_r, s>>>(P, h->last_noise, h->replay, h->top_idx, n, states_out_dev);
  }
  BNV_CUDA(cudaGetLastError());
  h->launches += 2;
  return BNV_OK;
}

int bnv_mppi_merge_top(bnv_mppi* h, const float* cand_dev, int32_t num_candidates, int32_t row_stride, int32_t n,
                       float* states_out_dev, float* weights_out_dev, void* stream) {
  if (!h || !cand_dev || !states_out_dev || !weights_out_dev) return fail(BNV_ERR_INVALID, "null argument");
  const int row_len = 3 * (h->P.T + 1);
  if (row_stride < row_len + 1) return fail(BNV_ERR_INVALID, "row_stride %d too small for a weight and %d state words", row_stride, row_len);
  if (n < 1 || n > num_candidates) return fail(BNV_ERR_INVALID, "n %d outside [1, %d]", n, num_candidates);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  // candidate weights gathered into a dense scratch (the selection kernel reads a contiguous array), then the same
  // radix select + sort as get_top_samples, then a row gather out of the candidate table
  const size_t need = static_cast<size_t>(num_candidates);
  if (h->merge_w_cap < need) {
    BNV_CUDA(cudaStreamSynchronize(s));
    if (h->merge_w) BNV_CUDA(cudaFree(h->merge_w));
    h->merge_w = nullptr;
    BNV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->merge_w), sizeof(float) * need));
    h->merge_w_cap = need;
  }
  if (h->top_idx_cap < static_cast<size_t>(n)) {
    BNV_CUDA(cudaStreamSynchronize(s));
    if (h->top_idx) BNV_CUDA(cudaFree(h->top_idx));
    h->top_idx = nullptr;
    BNV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->top_idx), sizeof(int) * n));
    h->top_idx_cap = n;
  }
  BNV_CUDA(cudaMemcpy2DAsync(h->merge_w, sizeof(float), cand_dev, sizeof(float) * row_stride, sizeof(float), num_candidates,
                             cudaMemcpyDeviceToDevice, s));
  int n_pad = 1;
  while (n_pad < n) n_pad <<= 1;
  unsigned long long* pairs_global = nullptr;
  size_t smem = static_cast<size_t>(n_pad) * 8;
  if (n_pad > bnv::kTopnSmemPairs) {
    if (h->top_pairs_cap < static_cast<size_t>(n_pad)) {
      BNV_CUDA(cudaStreamSynchronize(s));
      if (h->top_pairs) BNV_CUDA(cudaFree(h->top_pairs));
      h->top_pairs = nullptr;
      BNV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->top_pairs), sizeof(unsigned long long) * n_pad));
      h->top_pairs_cap = n_pad;
    }
    pairs_global = h->top_pairs;
    smem = 0;
  }
  BNV_CUDA(cudaFuncSetAttribute(bnv::topn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bnv::kTopnSmemPairs * 8));
  bnv::topn_select_kernel<<<1, bnv::kTopnThreads, smem, s>>>(h->merge_w, num_candidates, n, n_pad, pairs_global,
                                                             weights_out_dev, h->top_idx, nullptr, 0, nullptr);
  BNV_CUDA(cudaGetLastError());
  bnv::gather_strided_rows_kernel<<<n, 128, 0, s>>>(cand_dev + 1, h->top_idx, row_len, row_stride, states_out_dev);
  BNV_CUDA(cudaGetLastError());
  h->launches += 2;
  return BNV_OK;
}

float* bnv_mppi_weights(bnv_mppi* h) { return h ? h->weights : nullptr; }
float* bnv_mppi_costs(bnv_mppi* h) { return h ? h->costs : nullptr; }
float* bnv_mppi_states(bnv_mppi* h) { return h ? h->rec : nullptr; }
float* bnv_mppi_noise(bnv_mppi* h) { return h ? h->noise : nullptr; }
float* bnv_mppi_u_prev(bnv_mppi* h) { return h ? h->u_prev : nullptr; }
int32_t bnv_mppi_local_samples(const bnv_mppi* h) { return h ? h->Kl : 0; }
int32_t bnv_mppi_sample_offset(const bnv_mppi* h) { return h ? h->k_offset : 0; }

int bnv_mppi_reset(bnv_mppi* h, void* stream) {
  if (!h) return fail(BNV_ERR_INVALID, "null argument");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  BNV_CUDA(cudaMemsetAsync(h->u_prev, 0, sizeof(float) * h->E * h->P.T * 2, static_cast<cudaStream_t>(stream)));
  h->iteration = 0;
  if (h->iter_dev) BNV_CUDA(cudaMemsetAsync(h->iter_dev, 0, sizeof(unsigned long long), static_cast<cudaStream_t>(stream)));
  h->have_weights = false;
  return BNV_OK;
}

int bnv_mppi_draw_noise(bnv_mppi* h, uint64_t iteration, void* stream) {
  if (!h) return fail(BNV_ERR_INVALID, "null argument");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  const int T = h->P.T;
  const long long work = static_cast<long long>(h->Kl) * ((T + 1) / 2);
  const int blocks = static_cast<int>((work + 255) / 256);
  bnv::noise_kernel<<<dim3(blocks, h->E), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      h->noise, h->Kl, T, h->k_offset, static_cast<uint32_t>(h->cfg.seed), static_cast<uint32_t>(h->cfg.seed >> 32),
      static_cast<uint32_t>(iteration), static_cast<uint32_t>(iteration >> 32), h->cfg.sigma[0], h->cfg.sigma[1]);
  BNV_CUDA(cudaGetLastError());
  h->launches++;
  return BNV_OK;
}

int bnv_mppi_draw_xi(bnv_mppi* h, uint64_t iteration, float* xi_out_dev, float* xi_opt_out_dev, void* stream) {
  if (!h || !xi_out_dev || !xi_opt_out_dev) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->stoch) return fail(BNV_ERR_STATE, "not a stochastic-slip solver");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  const int T = h->P.T;
  const long long work = static_cast<long long>(h->Kl + 1) * ((T + 1) / 2);
  const int blocks = static_cast<int>((work + 255) / 256);
  bnv::xi_kernel<<<dim3(blocks, h->E), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      xi_out_dev, xi_opt_out_dev, h->Kl, T, h->k_offset, static_cast<uint32_t>(h->cfg.seed),
      static_cast<uint32_t>(h->cfg.seed >> 32), static_cast<uint32_t>(iteration), static_cast<uint32_t>(iteration >> 32));
  BNV_CUDA(cudaGetLastError());
  h->launches++;
  return BNV_OK;
}

int bnv_mppi_device_counter(bnv_mppi* h, int32_t enable, void* stream) {
  if (!h) return fail(BNV_ERR_INVALID, "null argument");
  if (h->cfg.world_size != 1) return fail(BNV_ERR_INVALID, "graph-capturable launches need world_size == 1");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  if (enable) {
    if (!h->iter_dev) BNV_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->iter_dev), sizeof(unsigned long long)));
    const unsigned long long it = h->iteration;
    BNV_CUDA(cudaMemcpyAsync(h->iter_dev, &it, sizeof(it), cudaMemcpyHostToDevice, s));
    // the "merge done" flags hold by-value epochs of earlier launches: clear them (0 is never a valid epoch)
    BNV_CUDA(cudaMemsetAsync(h->ticket, 0, 2 * static_cast<size_t>(h->E) * sizeof(unsigned int), s));
    BNV_CUDA(cudaStreamSynchronize(s));
    h->P.iter_dev = h->iter_dev;
    h->iter_external = enable == 2;
  } else if (h->iter_dev && h->P.iter_dev) {
    BNV_CUDA(cudaMemsetAsync(h->ticket, 0, 2 * static_cast<size_t>(h->E) * sizeof(unsigned int), s));
    unsigned long long it = 0;
    BNV_CUDA(cudaMemcpyAsync(&it, h->iter_dev, sizeof(it), cudaMemcpyDeviceToHost, s));
    BNV_CUDA(cudaStreamSynchronize(s));
    h->iteration = it;
    h->P.iter_dev = nullptr;
  }
  return BNV_OK;
}

uint64_t* bnv_mppi_iteration_counter(bnv_mppi* h) {
  return (h && h->P.iter_dev) ? reinterpret_cast<uint64_t*>(h->iter_dev) : nullptr;
}

int bnv_mppi_set_keep_mean(bnv_mppi* h, int32_t keep) {
  if (!h) return fail(BNV_ERR_INVALID, "null argument");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);  // a pre-launched kernel carries a by-value copy of the old parameters
  h->P.keep_mean = keep ? 1 : 0;
  return BNV_OK;
}

int bnv_mppi_set_terminal_goal(bnv_mppi* h, const float goal_xy[2]) {
  if (!h || !goal_xy) return fail(BNV_ERR_INVALID, "null argument");
  if (h->E > 1) return fail(BNV_ERR_INVALID, "not available on a batched solver");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  h->P.term_gx = goal_xy[0];
  h->P.term_gy = goal_xy[1];
  return BNV_OK;
}

int bnv_mppi_set_goal_dev(bnv_mppi* h, const float* goal_dev) {
  if (!h) return fail(BNV_ERR_INVALID, "null argument");
  if (h->E > 1) return fail(BNV_ERR_INVALID, "not available on a batched solver");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  h->P.goals = goal_dev;
  return BNV_OK;
}

int bnv_mppi_argmin(bnv_mppi* h, const float* actions_dev, float* action_out_dev, float* states_out_dev,
                    int32_t* index_out_dev, void* stream) {
  if (!h || !actions_dev || !action_out_dev || !states_out_dev) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->P.record) return fail(BNV_ERR_STATE, "argmin needs BNV_FLAG_RECORD_STATES");
  if (!h->have_weights || h->E > 1) return fail(BNV_ERR_STATE, "argmin needs a completed forward of a single solver");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  int rc = bnv_launch_argmin(h->costs, h->Kl, actions_dev, h->rec, 3 * (h->P.T + 1), action_out_dev, states_out_dev,
                             index_out_dev, static_cast<cudaStream_t>(stream));
  if (rc == BNV_OK) h->launches++;
  return rc;
}

int bnv_mppi_dwa_subgoal(bnv_mppi* h, const float* path_dev, int32_t n, const float* state_dev, const float* actions_dev,
                         float lookahead_distance, float* goal_out_dev, void* stream) {
  if (!h || !path_dev || !state_dev || !actions_dev || !goal_out_dev) return fail(BNV_ERR_INVALID, "null argument");
  if (n < 1) return fail(BNV_ERR_INVALID, "empty reference path");
  if (!h->problem_set) return fail(BNV_ERR_STATE, "bnv_mppi_set_problem must be called first");
  if (h->stoch || h->E > 1) return fail(BNV_ERR_INVALID, "needs a deterministic single solver");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  BNV_USER_WORK(h, stream);
  int rc = bnv_launch_dwa_subgoal(h->P.geom, h->P.G, h->tau, h->P.pitch, h->P.bounds, actions_dev, path_dev, n, state_dev,
                                  lookahead_distance, goal_out_dev, static_cast<cudaStream_t>(stream));
  if (rc == BNV_OK) h->launches++;
  return rc;
}

uint64_t bnv_mppi_launch_count(const bnv_mppi* h) { return h ? h->launches : 0; }

int bnv_mppi_check(bnv_mppi* h, void* stream) {
  if (!h) return fail(BNV_ERR_INVALID, "null argument");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  unsigned int flag = 0;
  BNV_CUDA(cudaMemcpyAsync(&flag, h->err_flag, sizeof(flag), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  BNV_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  if (flag != 0u) {
    BNV_CUDA(cudaMemsetAsync(h->err_flag, 0, sizeof(flag), static_cast<cudaStream_t>(stream)));
    return fail(BNV_ERR_CUDA, "an in-kernel wait timed out (a peer rank did not deliver its softmax partial within 2 s): "
                              "the results of that iteration are invalid");
  }
  return BNV_OK;
}

int bnv_mppi_launch_geometry(const bnv_mppi* h, int32_t out[4]) {
  if (!h || !out) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->problem_set) return fail(BNV_ERR_STATE, "bnv_mppi_set_problem must be called first");
  const bool coop = h->coop_ok && static_cast<long long>(h->grid) * h->E <= h->resident_ctas;
  out[0] = h->grid;
  out[1] = h->warps;
  out[2] = h->P.rec_split;
  out[3] = coop ? 1 : 0;
  return BNV_OK;
}

int bnv_mppi_kernel_timing(bnv_mppi* h, int32_t max_launches) {
  if (!h || max_launches < 0) return fail(BNV_ERR_INVALID, "bad argument");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);
  while (h->ev.size() < 2 * static_cast<size_t>(max_launches)) {
    cudaEvent_t e;
    BNV_CUDA(cudaEventCreate(&e));
    h->ev.push_back(e);
  }
  h->timing = max_launches > 0;
  h->ev_used = 0;
  return BNV_OK;
}

int bnv_mppi_kernel_time(bnv_mppi* h, double* total_ms, uint64_t* launches) {
  if (!h || !total_ms || !launches) return fail(BNV_ERR_INVALID, "null argument");
  BNV_CUDA(cudaSetDevice(h->cfg.device));
  BNV_DRAIN(h);  // else the event of a launch still waiting for its state would block until the time-out
  double sum = 0.0;
  for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
    float ms = 0.0f;
    BNV_CUDA(cudaEventSynchronize(h->ev[i + 1]));
    BNV_CUDA(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
    sum += ms;
  }
  *total_ms = sum;
  *launches = h->ev_used / 2;
  h->ev_used = 0;
  return BNV_OK;
}

int bnv_debug_timestamps(bnv_mppi* h, long long out[24]) {
  if (!h || !out) return fail(BNV_ERR_INVALID, "null argument");
  if (!h->dbg_ts) return fail(BNV_ERR_STATE, "set BNV_DEBUG_TS=1 before creating the handle");
  BNV_CUDA(cudaMemcpy(out, h->dbg_ts, 24 * sizeof(long long), cudaMemcpyDeviceToHost));
  return BNV_OK;
}

int bnv_debug_flush(void* buf_dev, uint64_t bytes, uint32_t smem_bytes, uint32_t value, void* stream) {
  if (!buf_dev || bytes < 16) return fail(BNV_ERR_INVALID, "bad argument");
  const int read_only = (smem_bytes >> 31) & 1u;  // top bit of smem_bytes: read the buffer instead of writing it
  smem_bytes &= 0x7FFFFFFFu;
  if (smem_bytes > kMaxDynSmem) return fail(BNV_ERR_INVALID, "too much shared memory");
  BNV_CUDA(cudaFuncSetAttribute(bnv::flush_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaxDynSmem)));
  bnv::flush_debug_kernel<<<148 * 4, 256, smem_bytes, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint4*>(buf_dev), bytes / 16, value, read_only);
  BNV_CUDA(cudaGetLastError());
  return BNV_OK;
}

int bnv_debug_philox(const uint32_t* in_dev, uint32_t* out_dev, int32_t n, void* stream) {
  if (!in_dev || !out_dev || n < 0) return fail(BNV_ERR_INVALID, "bad argument");
  if (n == 0) return BNV_OK;
  bnv::philox_debug_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(in_dev, out_dev, n);
  BNV_CUDA(cudaGetLastError());
  return BNV_OK;
}

int bnv_debug_sincos(const float* theta_dev, float* sin_dev, float* cos_dev, int32_t n, void* stream) {
  if (!theta_dev || !sin_dev || !cos_dev || n < 0) return fail(BNV_ERR_INVALID, "bad argument");
  if (n == 0) return BNV_OK;
  bnv::sincos_debug_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(theta_dev, sin_dev, cos_dev, n);
  BNV_CUDA(cudaGetLastError());
  return BNV_OK;
}

}  // extern "C"
