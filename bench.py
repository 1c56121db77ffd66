#!/usr/bin/env python
"""bench.py -- MPPI control iterations/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): synthetic 256x256 terrain (r = 0.5 m), K = 16384 samples,
T = 50 steps, sigma = (0.5, 0.5), lambda = 0.5, start (8, 8, pi/4), goal (48, 48), threshold 0.3, full contract
(noise drawn in-engine and kept, recorded states and weights written).  A "step" is one MPPI.forward().
At N > 1 GPUs the job is ONE solver whose samples are sharded (K = 16384 per GPU, weak scaling; 512x512 terrain
as in configs[2]) with one all-gather of the (m, s, U) softmax partial per step.

Numbers
  value        device-timed: per-step CUDA-event pairs around forward() (state resident in HBM), L2 flushed
               between steps by writing a 256 MiB buffer, max over ranks; unit = iterations of one
               (16384-sample x 50-step) shard per second summed over ranks (= control iterations/s x N).
  e2e          forward_host(state, out=...): every step the 12-byte state goes host -> device in the kernel's launch
               packet, the iteration runs, the kernel stores u* and the optimal state sequence (1012 bytes) device ->
               host into pinned mapped memory and raises a completion word the host polls; wall clock per step (L2
               flushed between steps, flush not counted).
  roofline     rollout kernel alone: SURVEY 8d algorithmic bytes / its CUDA-event duration (second pass with the
               engine's kernel-event recorder on), against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline the oracle port of the reference loop (oracle/mppi_oracle.py, PyTorch CPU ops) on the host cores.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_PER_GPU, HORIZON, SIGMAS, LAMBDA, RESOLUTION, SEED = 16384, 50, (0.5, 0.5), 0.5, 0.5, 42
FLUSH_BYTES = 256 << 20
METRIC = "mppi_iters_per_sec"
UNIT = "iters/s"


def algorithmic_bytes(k: int, t: int, g: int, channels: int = 1) -> int:
    """noise 8KT + recorded states 12K(T+1) + weights 4K + map 4CG^2 + outputs 8T + 12(T+1)."""
    return 8 * k * t + 12 * k * (t + 1) + 4 * k + 4 * channels * g * g + 8 * t + 12 * (t + 1)


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload: str):
    """Per-launch DRAM bytes of the rollout kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "rollout_traffic.json")) as f:
            d = json.load(f)
        return d.get(workload, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_problem(grid: int):
    import torch  # noqa: F401

    from benchnav_b200.synthetic import benchmark_problem

    return benchmark_problem(grid, RESOLUTION, seed=0)


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_rate(grid: int, k: int, steps: int, warmup: int, budget_s: float = 120.0):
    """The oracle port of the reference loop on the host cores.  Each step is one forward() of the full workload
    when `steps + warmup` of them fit the time budget; otherwise each step is a bounded sample -- all K samples
    over the first T_s < T horizon steps (and, below T_s = 1, fewer samples) -- and the rate is scaled linearly by
    the sampled fraction of the K x T rollout steps (the reference loop is a Python loop over T of elementwise
    [K]-wide ATen ops, so its time is proportional to both)."""
    import torch

    from oracle import mppi_oracle as orc

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    risk, start, goal, thr = build_problem(grid)
    p = orc.make_problem(risk, RESOLUTION, goal.tolist(), thr)
    probe = orc.OracleSolver(p, HORIZON, k, SIGMAS, LAMBDA, seed=SEED)
    probe.forward(start)
    t0 = time.perf_counter()
    probe.forward(start)
    est = time.perf_counter() - t0
    frac = budget_s / (est * (steps + warmup))
    k_run, t_run = k, HORIZON
    if frac < 1.0:
        t_run = max(1, int(HORIZON * frac))
        if HORIZON * frac < 1.0:
            k_run = max(1024, int(k * frac * HORIZON))
    scale = (k_run * t_run) / (k * HORIZON)
    sample = f"{steps} forward() calls of K={k_run}, T={t_run} on G={grid} after {warmup} warm-up"
    sample += " (full workload)" if scale == 1.0 else f"; rate scaled by {k_run}*{t_run}/({k}*{HORIZON}) to the full workload"
    solver = orc.OracleSolver(p, t_run, k_run, SIGMAS, LAMBDA, seed=SEED)
    for _ in range(warmup):
        solver.forward(start)
    t0 = time.perf_counter()
    for _ in range(steps):
        solver.forward(start)
    dt = time.perf_counter() - t0
    return steps / dt * scale, dt / steps * 1e3 / scale, cores, sample


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    grid = 256 if args.gpus == 1 else 512
    k = K_PER_GPU * args.gpus
    rate, ms, cores, sample = cpu_reference_rate(grid, k, args.steps, max(args.warmup, 1))
    value = rate * args.gpus  # same unit as the native arm: 16384-sample shard iterations per second
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"G={grid} terrain, K={k}, T={HORIZON}, oracle port of the reference PyTorch loop on CPU",
                   "grid": grid, "num_samples": k, "horizon": HORIZON},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ native arm
def run_native(args) -> None:
    import torch
    import torch.distributed as dist

    from benchnav_b200 import MPPI
    from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    grid = 256 if world == 1 else 512
    k_total = K_PER_GPU * world
    risk, start, goal, thr = build_problem(grid)
    dyn = UnicycleProblem(GridSpec(grid, RESOLUTION), risk)
    obj = GoalObjectives(dyn, goal, thr)
    solver = MPPI(HORIZON, k_total, 3, 2, dyn, obj, torch.tensor(SIGMAS), LAMBDA, device=dev, seed=SEED,
                  process_group=group, exchange=args.exchange)
    state_dev = start.to(dev)
    state_pinned = start.clone().pin_memory()
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    steps, warmup = args.steps, max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_pass(n: int):
        """n forward() calls, each preceded by an L2 flush, each bracketed by its own CUDA-event pair."""
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        barrier()
        for i in range(n):
            flush.fill_(i & 0xFF)
            ev0[i].record()
            solver.forward(state_dev)
            ev1[i].record()
        barrier()
        return sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))  # milliseconds

    for _ in range(warmup):
        flush.fill_(1)
        solver.forward(state_dev)
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = solver.launch_count
    total_ms = timed_pass(steps)
    launches = solver.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # second pass: rollout kernel alone (engine-side event pairs) -> roofline
    n_k = min(steps, 4096)
    solver.kernel_timing(n_k)
    timed_pass(n_k)
    kern_ms, kern_n = solver.kernel_time()
    solver.kernel_timing(0)
    kern_s = kern_ms / max(kern_n, 1) * 1e-3

    # back-to-back (no flush): explains how far launch gaps / cold L2 matter
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        solver.forward(state_dev)
    e1.record()
    barrier()
    hot_ms = e0.elapsed_time(e1) / steps

    # end to end through the host-buffer call (single GPU): pinned state in, results out, sync, every step
    e2e = None
    if world == 1:
        n_e = min(steps, 2000)
        out_bufs = (torch.empty(HORIZON, 2).pin_memory(), torch.empty(1, HORIZON + 1, 3).pin_memory())
        cur = torch.cuda.current_stream(dev)

        def e2e_pass(n: int) -> float:
            """n steps: L2 flush (not timed; only the flush's own stream is synchronised -- a pre-launched kernel is
            meant to be waiting on the device at that point), then the timed host-buffer call."""
            for _ in range(3):
                solver.forward_host(state_pinned, out=out_bufs)
            total = 0.0
            for i in range(n):
                flush.fill_(i & 0xFF)
                cur.synchronize()
                t0 = time.perf_counter()
                solver.forward_host(state_pinned, out=out_bufs)
                total += time.perf_counter() - t0
            return total

        acc_plain = e2e_pass(n_e)
        # pre-launched iterations: the next kernel is already resident when the state arrives (bnv_mppi_prelaunch)
        e2e_mode = "pre-launched iterations"
        try:
            solver.prelaunch(True)
            acc = e2e_pass(n_e)
        except Exception as exc:  # keep the bench line: report the plain-launch path as the end-to-end number
            acc, e2e_mode = acc_plain, f"plain launches (pre-launching failed: {exc})"
        finally:
            try:
                solver.prelaunch(False)
            except Exception:
                pass
        # the reference-style call that allocates fresh result tensors every step, for comparison
        acc_alloc = 0.0
        for i in range(min(n_e, 500)):
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            solver.forward_host(state_pinned)
            acc_alloc += time.perf_counter() - t0
        e2e = {"value": n_e / acc, "unit": UNIT, "h2d_bytes_per_step": 12,
               "d2h_bytes_per_step": 4 * (2 * HORIZON + 3 * (HORIZON + 1)), "steps": n_e,
               "mode": e2e_mode, "value_plain_launch": n_e / acc_plain,
               "value_allocating_outputs": min(n_e, 500) / acc_alloc,
               "timing": "wall clock around forward_host(state, out=caller buffers) per step with pre-launched "
                         "iterations (solver.prelaunch(): the 12-byte state is posted to a pinned, device-mapped mailbox "
                         "that the already-resident kernel polls; results D2H by zero-copy stores to pinned host memory, "
                         "host polls the kernel's completion word; the launch of the next iteration is issued inside the "
                         "timed call), L2 flushed before each step; value_plain_launch = the same call launching the "
                         "kernel when the state arrives (state in the launch packet); value_allocating_outputs = plain "
                         "launch, allocating fresh result tensors every step"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / steps
    control_rate = 1e3 / ms_per_step
    peak, peak_src = hbm_peak()
    alg_bytes = algorithmic_bytes(K_PER_GPU, HORIZON, grid)
    achieved = alg_bytes / kern_s / 1e9 if kern_s > 0 else 0.0
    workload = f"G{grid}_K{K_PER_GPU}_T{HORIZON}"
    line = {
        "metric": METRIC, "value": control_rate * world, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: {grid}x{grid} synthetic terrain, K={K_PER_GPU} samples/GPU "
                               f"(total {k_total}), T={HORIZON}, full contract (in-engine Philox noise kept, recorded "
                               f"states + weights written)",
                   "grid": grid, "num_samples_total": k_total, "horizon": HORIZON, "parallelism": f"sample-shard x{world}" + ("" if world == 1 else
                                   (", fused P2P mailbox exchange in-kernel" if solver._fused_exchange else
                                    ", NCCL all-gather + finalize kernel")),
                   "l2": "flushed between timed steps (256 MiB write); per-step CUDA-event pairs",
                   "launch": solver.launch_geometry,
                   "control_iters_per_sec": control_rate, "back_to_back_ms_per_step": hot_ms,
                   "rollout_steps_per_sec": control_rate * k_total * HORIZON},
        "roofline": {"bound": "hbm", "kernel": "bnv::rollout_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(workload), "algorithmic_bytes": alg_bytes,
                     "kernel_us": kern_s * 1e6, "launches_timed": kern_n, "peak_source": peak_src,
                     "note": "latency-bound by the T-step dependency chain at ~1 warp per SM sub-partition (DESIGN.md)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if e2e is not None:
        line["e2e"] = e2e
        rate, ms, cores, sample = cpu_reference_rate(grid, K_PER_GPU, 5, 2)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    else:
        line["e2e"] = {"value": control_rate * world, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                       "note": "multi-GPU: device-resident state; host-buffer path is single-GPU"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", choices=("native", "reference"), default="native")
    ap.add_argument("--exchange", choices=("p2p", "nccl"), default="p2p",
                    help="multi-GPU softmax exchange: fused over NVLink peer memory (default) or NCCL all-gather")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
