/*
 * bnv_mppi.h -- C ABI of the B200-native MPPI engine (libbnvmppi.so, sm_100a).
 *
 * Drop-in boundary for ONE path of masafumiendo/benchnav: a control iteration of the MPPI local
 * planner.  The reference has no FFI (it is pure Python/PyTorch); the entry points below are what a
 * binding for that path would call, and each cites the reference code it replaces (paths relative to
 * the reference repository root).  The Python mirror of the reference class that sits on top of this
 * ABI is benchnav_b200/mppi.py; the ctypes stub is benchnav_b200/_cabi.py.
 *
 * Conventions
 *   - every function returns 0 (BNV_OK) or a negative bnv_status; nothing throws across the ABI;
 *     bnv_last_error() returns a thread-local, NUL-terminated description of the last failure.
 *   - "dev" pointers are device pointers on the handle's CUDA device; the CALLER owns every buffer it
 *     passes in, the handle owns its scratch (noise, recorded states, costs, weights, mean sequence).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All *_dev entry points
 *     are asynchronous on that stream and never synchronise with the host.
 *   - one handle per (device, host thread); handles are not re-entrant (same as the reference module).
 *   - all arithmetic is IEEE fp32 in the reference's evaluation order (no fast-math, no FMA contraction
 *     in the state update) -- see DESIGN.md "Parity arithmetic".
 */
#ifndef BNV_MPPI_H_
#define BNV_MPPI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNV_ABI_VERSION 3

typedef enum bnv_status {
  BNV_OK = 0,
  BNV_ERR_INVALID = -1,     /* bad argument (the reference would raise AssertionError/ValueError) */
  BNV_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed; see bnv_last_error() */
  BNV_ERR_UNSUPPORTED = -3, /* valid request this build cannot serve (e.g. horizon too long for shared memory) */
  BNV_ERR_STATE = -4        /* call order violated (forward before set_problem, top_samples before forward) */
} bnv_status;

/* bnv_mppi_cfg.flags */
#define BNV_FLAG_RECORD_STATES 0x1u /* keep every sample's recorded state sequence (reference `_state_seq_batch`,
                                       mppi.py:119-125); required by bnv_mppi_top_samples */
#define BNV_FLAG_STOCHASTIC_SLIP 0x2u /* BASELINE config 4: every traversability lookup of the rollouts draws
                                         1 - clamp(Normal(mean, std).sample(), 0, 1) from the cell's slip distribution
                                         (the observation-mode lookup, traversability_model.py:65-69, applied to
                                         GridMap.distributions["predictions"]) instead of reading the risk map;
                                         needs bnv_mppi_set_problem_ex with a std map; implies RECORD_STATES;
                                         world_size must be 1 */

/* Constructor arguments of the reference `MPPI.__init__` (src/planners/local_planners/mppi.py:23-36) plus
 * the action bounds it copies from `dynamics.min_action/max_action` (mppi.py:83-88, robot_model.py:54-57),
 * the rollout time step (`transit`'s default delta_t, robot_model.py:60) and the sample-shard geometry. */
typedef struct bnv_mppi_cfg {
  int32_t num_samples;    /* K: TOTAL samples over all shards (mppi.py:26) */
  int32_t horizon;        /* T (mppi.py:25) */
  float sigma[2];         /* noise std-dev per control dim (mppi.py:31) */
  float lambda_;          /* temperature (mppi.py:32) */
  float u_min[2];         /* robot_model.py:55 */
  float u_max[2];         /* robot_model.py:56 */
  float dt;               /* robot_model.py:60 */
  uint64_t seed;          /* mppi.py:35; keys the in-kernel Philox stream */
  int32_t rank;           /* this shard's index, 0 <= rank < world_size */
  int32_t world_size;     /* number of sample shards (GPUs); 1 = single GPU */
  int32_t device;         /* CUDA device ordinal */
  uint32_t flags;         /* BNV_FLAG_* */
  int32_t num_envs;       /* E: independent environments solved per forward call (BASELINE config 3: one planner per
                             planetary_env instance, all in one launch); 0 or 1 = a single solver.  With E > 1 every
                             per-solver buffer below gains a leading dimension E, num_samples is per environment,
                             world_size must be 1 (environments, not samples, are what shards across GPUs) and
                             RECORD_STATES is implied. */
} bnv_mppi_cfg;

typedef struct bnv_mppi bnv_mppi; /* opaque solver handle */

/* Library / ABI identification. */
int bnv_abi_version(void);
const char* bnv_last_error(void);

/* MPPI.__init__ (mppi.py:23-128): validates, allocates the shard's buffers (noise [Kl,T,2], recorded
 * states [Kl,T+1,3], costs/weights [Kl]) and zeroes the mean sequence (mppi.py:116).  Kl = this rank's
 * share of K: global samples [rank*K/W, (rank+1)*K/W). */
int bnv_mppi_create(bnv_mppi** out, const bnv_mppi_cfg* cfg);
void bnv_mppi_destroy(bnv_mppi* h);

/* What forward reads through `dynamics`/`objectives`: the risk map `TraversabilityModel._risks`
 * (traversability_model.py:28-51; [G,G] fp32, row = y cell, `pitch` elements between rows), the GridMap
 * geometry (grid_map.py:40-50) and `Objectives._goal_pos/_stuck_threshold` (objectives.py:25-27).
 * Builds the engine's traversability map 1 - clamp(risk,0,1) (traversability_model.py:71-72) and its
 * TMA descriptor.  Must be called again if the risk map's contents change. */
int bnv_mppi_set_problem(bnv_mppi* h, const float* risk_dev, int32_t grid_size, int32_t pitch, float resolution,
                         float x_min, float x_max, float y_min, float y_max, const float goal_xy[2],
                         float stuck_threshold, void* stream);

/* General form of bnv_mppi_set_problem.
 *   mean_dev   [E][G][pitch] risk maps (deterministic mode) or slip means (BNV_FLAG_STOCHASTIC_SLIP);
 *              `env_stride` elements between consecutive environments' maps (0 = all environments share one map)
 *   std_dev    slip standard deviations, same layout (stochastic mode only; NULL otherwise)
 *   goals_xy   HOST array [E][2], one goal per environment (objectives.py:25) */
int bnv_mppi_set_problem_ex(bnv_mppi* h, const float* mean_dev, const float* std_dev, int32_t grid_size, int32_t pitch,
                            int64_t env_stride, float resolution, float x_min, float x_max, float y_min, float y_max,
                            const float* goals_xy, float stuck_threshold, void* stream);

/* MPPI.forward (mppi.py:130-219), device buffers.
 *   state_dev      [3]        current state (x, y, theta)
 *   noise_dev      [Kl,T,2]   sigma-scaled control noise exactly as `_action_noises` (mppi.py:149-151) for
 *                             this shard, or NULL to draw it in-kernel (Philox4x32-10 keyed by seed, global
 *                             sample index and iteration count => independent of world_size)
 *   u_out_dev      [T,2]      optimal control sequence (also kept as the next call's mean, mppi.py:217)
 *   opt_states_dev [T+1,3]    recorded states of the optimal rollout (mppi.py:202-214)
 * With world_size > 1 this runs the shard-local phase only (rollouts, costs, shard partial) and leaves
 * u_out/opt_states untouched; exchange bnv_mppi_partial() across ranks and call bnv_mppi_finalize(). */
int bnv_mppi_forward(bnv_mppi* h, const float* state_dev, const float* noise_dev, float* u_out_dev,
                     float* opt_states_dev, void* stream);

/* forward with injected lookup normals (BNV_FLAG_STOCHASTIC_SLIP; parity tests):
 *   xi_dev      [E][Kl][2T+1]  per sample: (transit lookup of step 0, stage-cost lookup of recorded state 0, transit 1,
 *                              stage 1, ..., terminal-cost lookup) -- the standard normals behind each
 *                              Normal(mean, std).sample() of robot_model.py:76 / objectives.py:50
 *   xi_opt_dev  [E][T]         transit lookups of the batch-1 optimal rollout (mppi.py:209-214)
 * noise_dev, xi_dev and xi_opt_dev must be all given or all NULL (NULL: drawn in-kernel from the Philox stream).
 * In batch mode (num_envs = E > 1) state_dev is [E][3], noise_dev [E][K][T][2], u_out_dev [E][T][2] and
 * opt_states_dev [E][T+1][3] -- for bnv_mppi_forward as well. */
int bnv_mppi_forward_ex(bnv_mppi* h, const float* state_dev, const float* noise_dev, const float* xi_dev,
                        const float* xi_opt_dev, float* u_out_dev, float* opt_states_dev, void* stream);

/* forward with the state given as three HOST floats (passed by value in the launch packet: no device copy of
 * the state is needed) and DEVICE outputs; asynchronous like bnv_mppi_forward. */
int bnv_mppi_forward_state(bnv_mppi* h, const float state_host[3], const float* noise_dev, float* u_out_dev,
                           float* opt_states_dev, void* stream);

/* Same call with HOST buffers (the reference-facing form: `forward(state)` takes and returns tensors the
 * caller reads on the host).  Host->device: the state rides in the launch packet.  Device->host: the kernel
 * stores u_out/opt_states into pinned, device-mapped staging (zero-copy over PCIe) and raises a completion word
 * there as soon as both are written; the call polls that word (falling back to a stream synchronisation if the
 * device faults or stalls) and copies the results to the caller's buffers.
 * See bnv_mppi_prelaunch for the variant in which the kernel is already resident when the state arrives.
 * Sharded solver (world_size > 1, peers attached): the control loop lives on ONE rank, the leader, which makes this
 * call; its kernel broadcasts the state to the other ranks' mailboxes over NVLink.  Every other rank calls
 * bnv_mppi_forward_follow once per leader call: its kernel waits on the device for the state, rolls out its shard
 * and takes part in the exchange; the leader's host buffers receive the (identical on every rank) results. */
int bnv_mppi_forward_host(bnv_mppi* h, const float state_host[3], const float* noise_dev, float* u_out_host,
                          float* opt_states_host, void* stream);
int bnv_mppi_forward_follow(bnv_mppi* h, const float* noise_dev, float* u_out_dev, float* opt_states_dev, void* stream);

/* Two-stage form of bnv_mppi_forward_host for a control loop that needs the controls first (the reference's loop
 * steps the environment with u*[0] and only draws the optimal trajectory, test/test_mppi.py:171-198).  The kernel
 * raises a first completion word when u* is in the staging buffer -- before the serial optimal rollout
 * (mppi.py:205-213), ~3.7 us at T = 50 -- and _action returns on it; bnv_mppi_wait_states then returns the optimal state
 * sequence [T+1][3] of that same launch (usually complete by the time it is asked for).  The staging buffer is reused
 * by the next launch: collect the states BEFORE the next forward_host / forward_host_action call, or not at all. */
/* Host-buffer form for a batched solver (num_envs = E >= 1): states_host [E][3] in, u_out_host [E][T][2] and
 * opt_states_host [E][T+1][3] out; the states go up in one staged copy, the results are stored by the kernel into the
 * handle's pinned, device-mapped buffer, one stream synchronisation (E x `forward(state)` + `.cpu()` of the reference's loop, mppi.py:130-219, in one call). */
int bnv_mppi_forward_host_batch(bnv_mppi* h, const float* states_host, float* u_out_host, float* opt_states_host,
                                void* stream);
int bnv_mppi_forward_host_action(bnv_mppi* h, const float state_host[3], float* u_out_host, void* stream);
int bnv_mppi_wait_states(bnv_mppi* h, float* opt_states_host);

/* Sample-sharded softmax (replaces the global torch.softmax of mppi.py:193-199).
 * bnv_mppi_partial: device pointer to this shard's (m, s, U[T,2]) -- m = max_k(-c_k/lambda),
 * s = sum_k exp(-c_k/lambda - m), U = sum_k exp(..) v[k] -- bnv_mppi_partial_len() floats, valid after
 * bnv_mppi_forward on `stream`.  bnv_mppi_finalize: log-sum-exp merge of `world_size` gathered partials
 * (rank-major), normalises this shard's weights, writes u_out/opt_states and the next mean sequence. */
const float* bnv_mppi_partial(const bnv_mppi* h);
int32_t bnv_mppi_partial_len(const bnv_mppi* h);
int bnv_mppi_finalize(bnv_mppi* h, const float* gathered_partials_dev, float* u_out_dev, float* opt_states_dev,
                      void* stream);

/* Fused exchange over NVLink peer memory (replaces the host-side all-gather + bnv_mppi_finalize pair): every rank
 * exports the CUDA IPC handle of its mailbox (64 bytes), the handles are gathered by the caller (any transport,
 * rank-major) and attached once.  Afterwards bnv_mppi_forward on a sharded handle exchanges the shard partials
 * inside the rollout kernel -- every column of the partial travels as self-validating 8-byte {value, sequence} words
 * stored straight into the peers' mailboxes, double-buffered by the parity of the sequence number; no fence, no flag,
 * no NCCL call -- and writes u_out/opt_states itself; all ranks must call forward the same number of times (as with
 * any collective).  A peer that does not deliver within 2 s makes the waiting kernel give up instead of hanging the
 * device; bnv_mppi_check() then reports the failure. */
int bnv_mppi_mailbox_handle(bnv_mppi* h, unsigned char out[64]);
int bnv_mppi_attach_peers(bnv_mppi* h, const unsigned char* handles /* [world_size][64] */);

/* MPPI.get_top_samples (mppi.py:221-240): the n highest-weight samples of this shard in descending
 * weight order.  states_out_dev [n,T+1,3], weights_out_dev [n].  With BNV_FLAG_RECORD_STATES the rows are gathered from
 * the recorded states; without it (a lean solver) the n selected samples are rolled out again from the iteration's
 * saved start state, mean sequence and noise rows -- bit-identical rows (if the last forward injected its noise, that
 * array must still be alive). */
int bnv_mppi_top_samples(bnv_mppi* h, int32_t n, float* states_out_dev, float* weights_out_dev, void* stream);
/* (batch mode: the n best samples of every environment, states_out_dev [E,n,T+1,3], weights_out_dev [E,n]) */

/* Sharded solver: global top-n out of the ranks' local top lists.  cand_dev [num_candidates][row_stride] holds one
 * candidate per row, {weight, states [T+1,3] ...} (row_stride >= 1 + 3 (T+1); the caller gathers the ranks' lists into
 * it with whatever transport it uses -- one all-gather).  Selects the n largest weights with the same radix select +
 * sort as bnv_mppi_top_samples and gathers their rows: states_out_dev [n,T+1,3], weights_out_dev [n], descending. */
int bnv_mppi_merge_top(bnv_mppi* h, const float* cand_dev, int32_t num_candidates, int32_t row_stride, int32_t n,
                       float* states_out_dev, float* weights_out_dev, void* stream);

/* Module state the reference exposes as attributes (device pointers owned by the handle, shard-local):
 *   weights  [Kl]        `_weights`            (mppi.py:193)
 *   costs    [Kl]        per-sample cost        (mppi.py:186-190; not stored by the reference)
 *   states   [Kl,T+1,3]  `_state_seq_batch`    (mppi.py:160-165), NULL without BNV_FLAG_RECORD_STATES
 *   noise    [Kl,T,2]    `_action_noises`      (mppi.py:149-151) of the last in-kernel draw
 *   u_prev   [T,2]       `_previous_action_seq` (mppi.py:116, :217) */
float* bnv_mppi_weights(bnv_mppi* h);
float* bnv_mppi_costs(bnv_mppi* h);
float* bnv_mppi_states(bnv_mppi* h);
float* bnv_mppi_noise(bnv_mppi* h);
float* bnv_mppi_u_prev(bnv_mppi* h);
int32_t bnv_mppi_local_samples(const bnv_mppi* h); /* Kl */
int32_t bnv_mppi_sample_offset(const bnv_mppi* h); /* first global sample index of this shard */

/* Zero the mean sequence and restart the noise stream (a fresh `MPPI(...)`, mppi.py:116). */
int bnv_mppi_reset(bnv_mppi* h, void* stream);

/* Fill the handle's noise buffer (bnv_mppi_noise) with the engine's Philox stream for the given iteration
 * index using the stand-alone noise kernel.  bnv_mppi_forward(noise_dev = NULL) draws the same values inside the
 * rollout kernel for iteration 0, 1, 2, ... since creation / bnv_mppi_reset; this entry point exists so that the
 * stream can be inspected or pre-drawn. */
int bnv_mppi_draw_noise(bnv_mppi* h, uint64_t iteration, void* stream);

/* The stochastic mode's lookup normals for the given iteration index, by the stand-alone kernel (the same Philox
 * calls the rollout kernel makes): xi_out_dev [E][Kl][2T+1], xi_opt_out_dev [E][T] (layouts of bnv_mppi_forward_ex),
 * caller-owned.  Lets a test replay an in-engine-noise iteration through the oracle. */
int bnv_mppi_draw_xi(bnv_mppi* h, uint64_t iteration, float* xi_out_dev, float* xi_opt_out_dev, void* stream);

/* Graph-capturable launches.  A captured CUDA graph freezes the kernel's launch packet, which normally carries the
 * iteration counter (the Philox counter word) and the launch epoch.  With enable != 0 both are read from a
 * device-resident counter instead, advanced by a one-thread kernel after every forward in stream order -- so a
 * stream capture of forward (+ the environment step, collision check, ...) can be replayed as one graph launch per
 * control step, with the same noise stream as uncaptured calls.  enable = 0 returns to by-value counters, keeping
 * the count.  Synchronises `stream`; world_size must be 1; bnv_mppi_forward_host is not capturable. */
int bnv_mppi_device_counter(bnv_mppi* h, int32_t enable, void* stream);
/* enable = 2: as 1, but the counter is advanced by the CALLER (bnv_closed_loop_step does it), not by a bump kernel
 * behind every forward.  bnv_mppi_iteration_counter: the device address of the counter (NULL unless enabled). */
uint64_t* bnv_mppi_iteration_counter(bnv_mppi* h);

/* Pre-launched iterations for the host-buffer call.  With enable != 0, bnv_mppi_forward_host queues the NEXT
 * iteration's kernel (on an internal stream) before it waits for the current one; that kernel becomes resident as soon
 * as the current one finishes and polls a host-mapped {state, sequence} word (one thread over PCIe, broadcast to the
 * other CTAs through device memory) -- so the following
 * bnv_mppi_forward_host call is a 16-byte store plus the completion poll: neither the launch call nor the launch
 * latency is on the step's critical path.  A pre-launched kernel whose state does not arrive within `timeout_us`
 * (0 = 2000) aborts itself (all CTAs follow one grid-wide decision) and the step falls back to a plain launch;
 * every other entry point cancels a waiting launch first.  Iteration numbers of cancelled / aborted launches are
 * given back, so results are identical to the plain path.  While a launch waits it occupies its SMs: work the caller
 * queues on other streams runs on the remaining ones, and a device-wide synchronisation lasts until the timeout.
 * Engine-owned tensors (weights, recorded states, ...) are ordered for the caller's stream by any other entry point. */
int bnv_mppi_prelaunch(bnv_mppi* h, int32_t enable, uint32_t timeout_us);

/* ---- hooks used by the DWA planner built on the same rollout kernel (src/planners/local_planners/dwa.py) ----
 * DWA.forward (dwa.py:116-149) = the MPPI rollout/cost machinery with K = num_lin_vel * num_ang_vel constant
 * action sequences injected as "noise" around a zero mean (bnv_mppi_forward with noise_dev = the held actions),
 * lambda_ = 1 (weights = softmax(-cost), dwa.py:147), argmin instead of the weighted mean.
 *   bnv_mppi_set_keep_mean(h, 0)      u* is not written back as the next mean sequence (the mean stays zero)
 *   bnv_mppi_set_terminal_goal        goal of the terminal cost (always the final goal, dwa.py:233) when the stage
 *                                     cost follows a sub-goal (dwa.py:225-231); reset by bnv_mppi_set_problem*
 *   bnv_mppi_set_goal_dev             device-resident [2] override of the stage-cost goal (the sub-goal selected by
 *                                     bnv_mppi_dwa_subgoal), NULL = back to the goal of bnv_mppi_set_problem
 *   bnv_mppi_argmin                   first index of the minimum cost of the last forward, that sample's action
 *                                     (from actions_dev [K,2]) and recorded states [T+1,3] (dwa.py:141-144) */
int bnv_mppi_set_keep_mean(bnv_mppi* h, int32_t keep);
int bnv_mppi_set_terminal_goal(bnv_mppi* h, const float goal_xy[2]);
int bnv_mppi_set_goal_dev(bnv_mppi* h, const float* goal_dev);
int bnv_mppi_argmin(bnv_mppi* h, const float* actions_dev, float* action_out_dev, float* states_out_dev,
                    int32_t* index_out_dev, void* stream);
/* DWA._select_sub_goal (dwa.py:260-285) as DWA._compute_costs calls it (dwa.py:225-228): on
 * state_seq_batch[0, 0, :] AFTER the simulation, i.e. on the raw successor that transit's in-place update left in
 * slot 0 of the first action's rollout -- recomputed here from state_dev [3], actions_dev[0] and the handle's
 * traversability map.  path_dev [n,2] -> goal_out_dev [2] (feed it to bnv_mppi_set_goal_dev). */
int bnv_mppi_dwa_subgoal(bnv_mppi* h, const float* path_dev, int32_t n, const float* state_dev, const float* actions_dev,
                         float lookahead_distance, float* goal_out_dev, void* stream);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
uint64_t bnv_mppi_launch_count(const bnv_mppi* h);

/* Synchronises `stream` and reports whether an in-kernel wait of an earlier iteration timed out (sharded solvers: a
 * peer rank never delivered its partial).  BNV_OK, or BNV_ERR_CUDA with the reason in bnv_last_error(); the error
 * state is cleared by the call. */
int bnv_mppi_check(bnv_mppi* h, void* stream);

/* Launch geometry chosen for the rollout kernel by the last bnv_mppi_set_problem*: out = {CTAs per environment,
 * warps per CTA, the step at which the recorded-state slab is flushed mid-loop (0 = one flush at the end),
 * 1 if the grid is co-resident and launched cooperatively}. */
int bnv_mppi_launch_geometry(const bnv_mppi* h, int32_t out[4]);

/* Measurement hook: record a CUDA-event pair around the rollout kernel of each of the next `max_launches`
 * forward calls (0 disables); bnv_mppi_kernel_time() synchronises on them, returns the summed kernel time and
 * the number of launches measured, and re-arms the recorder. */
int bnv_mppi_kernel_timing(bnv_mppi* h, int32_t max_launches);
int bnv_mppi_kernel_time(bnv_mppi* h, double* total_ms, uint64_t* launches);

/* Debug hook: clock64() stamps taken by the last CTA of the most recent rollout kernel (phase boundaries;
 * see mppi_kernels.cuh BNV_STAMP).  Only recorded when the handle was created with BNV_DEBUG_TS set. */
int bnv_debug_timestamps(bnv_mppi* h, long long out[24]);
/* Measurement aid: an L2-flushing fill of buf_dev[0, bytes) launched with `smem_bytes` of dynamic shared memory (to
 * test whether the shared-memory carve-out switch between kernels is part of the event-timed launch floor).  Top bit
 * of smem_bytes set: the buffer is READ instead (L2 left full of clean lines rather than dirty ones). */
int bnv_debug_flush(void* buf_dev, uint64_t bytes, uint32_t smem_bytes, uint32_t value, void* stream);

/* Test hook: the raw Philox4x32-10 block function behind the engine's noise stream, for known-answer tests
 * (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 kat_vectors).
 * in_dev [n][6] = (counter x, y, z, w, key lo, key hi) -> out_dev [n][4]. */
int bnv_debug_philox(const uint32_t* in_dev, uint32_t* out_dev, int32_t n, void* stream);

/* Test hook: evaluates the engine's in-range sin/cos (used by the state update in place of
 * torch.cos/torch.sin, robot_model.py:86-87) on n device floats. */
int bnv_debug_sincos(const float* theta_dev, float* sin_dev, float* cos_dev, int32_t n, void* stream);

/* ================================================================================================
 * Rows either side of the MPPI iteration: stateless entry points, device pointers on
 * the CURRENT CUDA device, asynchronous on `stream`.
 * ================================================================================================ */

/* GridMap geometry (grid_map.py:40-50) of a [grid_size, grid_size] map stored with `pitch` elements per row. */
typedef struct bnv_grid {
  int32_t grid_size;
  int32_t pitch;
  float resolution;
  float x_min, x_max, y_min, y_max;
} bnv_grid;

/* TraversabilityModel.get_traversability (traversability_model.py:53-72) at n positions (rows of `pos_stride`
 * floats starting with x, y -- e.g. states [B,P,3] flattened, pos_stride 3) and PlanetaryEnv.collision_check
 * (planetary_env.py:221-232).
 *   std_dev == NULL  inference mode: trav = 1 - clamp(mean[cell], 0, 1) with mean = the risk map
 *   std_dev != NULL  observation mode: trav = 1 - clamp(mean[cell] + std[cell] * xi, 0, 1); xi_dev [n] standard
 *                    normals, or NULL to draw them from Philox(seed, counter + *counter_dev); counter_dev is an
 *                    optional device-resident 64-bit draw counter (NULL = 0) that the caller advances in stream order,
 *                    so that the call can be captured in a CUDA graph and still draw fresh normals on every replay
 *   rows_per_env > 0 positions are [E][rows_per_env] and environment e reads the map at mean_dev + e * env_stride
 *   trav_out_dev [n] and/or stuck_out_dev [n] (uint8, trav <= stuck_threshold); either may be NULL. */
int bnv_trav_lookup(const bnv_grid* grid, const float* mean_dev, const float* std_dev, int64_t env_stride,
                    int64_t rows_per_env, const float* pos_dev, int64_t n, int32_t pos_stride, const float* xi_dev,
                    uint64_t seed, uint64_t counter, const uint64_t* counter_dev, float stuck_threshold,
                    float* trav_out_dev, uint8_t* stuck_out_dev, void* stream);

/* PlanetaryEnv.step (planetary_env.py:189-219) for E independent environments in one launch:
 * observation-mode transit of states_dev [E,3] (updated in place) under actions_dev [E,2], reward_out_dev [E] = the
 * traversability drawn for the step, terminated_out_dev [E] (uint8) = ||p - goal|| < goal_threshold.  Maps as in
 * bnv_trav_lookup (env_stride 0 = shared).  xi_dev [E] or NULL (Philox(seed, counter + *counter_dev)).  The elapsed-time /
 * truncation bookkeeping (two scalars) stays with the caller. */
int bnv_env_step(const bnv_grid* grid, const float* mean_dev, const float* std_dev, int64_t env_stride,
                 int32_t num_envs, float* states_dev, const float* actions_dev, const float* goals_dev,
                 const float* xi_dev, uint64_t seed, uint64_t counter, const uint64_t* counter_dev, const float u_min[2],
                 const float u_max[2], float delta_t, float goal_threshold, float* reward_out_dev,
                 uint8_t* terminated_out_dev, void* stream);

/* Everything Tutorial 3.3's loop does between two planner calls (test/test_mppi.py:180-185), for E environments in ONE
 * launch: PlanetaryEnv.step with the first planned action actions_dev[e][0] (robots whose done_dev[e] is set stop),
 * PlanetaryEnv.collision_check of the planned trajectory planned_dev [E,T+1,3] -> collisions_out_dev [E,T+1] (uint8),
 * and the loop's books: done_dev[e] |= terminated, steps_to_goal_dev[e] = step number of the first arrival,
 * *step_no_dev += 1, *counter_dev += 2 (the draws are those of bnv_env_step at counter c followed by
 * bnv_trav_lookup at counter 2^40 + c + 1: bit-identical to the separate calls), and, when planner_iteration_dev is
 * given (bnv_mppi_iteration_counter of a solver in externally-advanced mode), *planner_iteration_dev += 1.
 * ticket_dev: one zero-initialised uint32 owned by the caller.  With this a graph-captured control step is two kernels. */
int bnv_closed_loop_step(const bnv_grid* grid, const float* mean_dev, const float* std_dev, int64_t env_stride,
                         int32_t num_envs, int32_t horizon, float* states_dev, const float* actions_dev,
                         const float* planned_dev, const float* goals_dev, uint64_t seed, uint64_t* counter_dev,
                         const float u_min[2], const float u_max[2], float delta_t, float goal_threshold,
                         float stuck_threshold, float* reward_out_dev, uint8_t* terminated_out_dev,
                         uint8_t* collisions_out_dev, uint8_t* done_dev, int64_t* steps_to_goal_dev, int64_t* step_no_dev,
                         uint64_t* planner_iteration_dev, uint32_t* ticket_dev, void* stream);

/* TraversabilityModel._infer_risk_map (traversability_model.py:28-51) over n_cells cells.
 *   metric      0 expected value, 1 VaR, 2 CVaR (utils.py:18); confidence = ModelConfig.confidence_value
 *   method      BNV_RISK_CLOSED_FORM: mean + coef * std (exact for the Normal slip model);
 *               BNV_RISK_MONTE_CARLO: the reference's estimator -- num_samples draws per cell, torch.quantile
 *               (linear) and the mean of the tail above it -- on injected samples_dev [num_samples, n_cells]
 *               (exactly `distributions.sample((num_samples,))`) or, with samples_dev NULL, on Philox(seed) draws;
 *               samples_out_dev (optional, same shape) receives those draws. */
#define BNV_RISK_CLOSED_FORM 0
#define BNV_RISK_MONTE_CARLO 1
int bnv_risk_map(int32_t metric, float confidence, int32_t method, const float* mean_dev, const float* std_dev,
                 int64_t n_cells, const float* samples_dev, int32_t num_samples, uint64_t seed, float* risk_out_dev,
                 float* samples_out_dev, void* stream);

/* DWA._generate_actions (dwa.py:151-184): actions_out_dev [nv*nw, 2] = cartesian product of the linspaces over the
 * dynamic window around prev_action_dev [2]; controls_out_dev [nv*nw, T, 2] = each action held over the horizon
 * (the form bnv_mppi_forward takes as injected noise). */
int bnv_dwa_actions(const float* prev_action_dev, const float u_min[2], const float u_max[2], const float a_lim[2],
                    float delta_t, int32_t num_lin_vel, int32_t num_ang_vel, int32_t horizon, float* actions_out_dev,
                    float* controls_out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BNV_MPPI_H_ */
