#!/usr/bin/env python
"""bench.py -- MPPI control iterations/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config auto|c1|c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json `configs`; all: r = 0.5 m, sigma = (0.5, 0.5), lambda = 0.5, start (8, 8, pi/4),
goal (0.375 G r, 0.375 G r), threshold 0.3, full contract = noise drawn in-engine and kept, recorded states and
weights written).  A "step" is one control iteration = one MPPI.forward() of the whole job.
  c1    configs[1]: 256x256 terrain, K = 16384, T = 50, one GPU                       (the headline; `auto` at N = 1)
  c2w   configs[2] family, weak scaling: 512x512, K = 16384 per GPU, T = 50, samples sharded over the N GPUs with one
        exchange of the (m, s, U) softmax partial per step; at N = 8 this IS configs[2]     (`auto` at N > 1)
  c2    configs[2] itself at any N: 512x512, K = 131072 in total, T = 50, samples sharded over N GPUs (strong scaling;
        N = 1 is the single-GPU datum)
  c3    configs[3]: 64 environments x K = 4096, T = 30, own 64x64 map each, environments sharded over N GPUs, no
        exchange at all (strong scaling)
  c4    configs[4]: stochastic slip, 256x256, K = 32768, T = 50, one GPU

Numbers
  value        device-timed: per-step CUDA-event pairs around forward() (inputs resident in HBM), L2 flushed between
               steps by writing a 256 MiB buffer, max over ranks.  Unit: control iterations/s of the job, except c2w
               (weak scaling) where it is iterations of one (16384-sample x 50-step) shard per second summed over the
               ranks = control iterations/s x N, so that N = 1 is exactly BASELINE's metric; the raw control rate is
               always in detail.control_iters_per_sec.
  e2e          the same step through the host-buffer call: inputs from pinned host memory, results back in host memory,
               copies inside the timed region, wall clock per step.
  roofline     rollout kernel alone: SURVEY 8d algorithmic bytes of one launch / its CUDA-event duration (second pass,
               engine-side event pairs), against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline the oracle port of the reference loop (oracle/mppi_oracle.py, PyTorch CPU ops) on the host cores, on a
               bounded sample of the same workload.
Multi-GPU lines (c2w / c2 at N > 1) also carry: `parity_check` (every rank holds the bit-identical u*, and it equals the
unsharded solver's within 5e-6, checked on the first step), and detail.single_gpu_same_total_ms / speedup_vs_1gpu
(rank 0 alone timed on the SAME total workload with the same method).
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_PER_GPU, SIGMAS, LAMBDA, RESOLUTION, SEED = 16384, (0.5, 0.5), 0.5, 0.5, 42
FLUSH_BYTES = 256 << 20
METRIC = "mppi_iters_per_sec"
UNIT = "iters/s"


# ------------------------------------------------------------------------------------------ workloads
def resolve_workload(name: str, world: int) -> dict:
    """Sizes of the named workload at `world` GPUs.  kind: single (one solver, samples sharded when world > 1),
    batch (independent environments, sharded by environment) or stoch (stochastic-slip lookups)."""
    if name == "auto":
        name = "c1" if world == 1 else "c2w"
    if name == "c1":
        w = dict(kind="single", grid=256, k_total=K_PER_GPU, horizon=50, envs=1, scaling="weak",
                 label="BASELINE configs[1]: 256x256 synthetic terrain, K=16384, T=50, single B200")
        if world != 1:
            raise SystemExit("c1 is the single-GPU configuration; use c2w / c2 for sample-sharded runs")
    elif name == "c2w":
        w = dict(kind="single", grid=512, k_total=K_PER_GPU * world, horizon=50, envs=1, scaling="weak",
                 label=f"BASELINE configs[2] family (weak scaling): 512x512 synthetic terrain, K=16384 per GPU "
                       f"(total {K_PER_GPU * world}), T=50, samples sharded over {world} B200" +
                       (" -- exactly configs[2]" if world == 8 else ""))
    elif name == "c2":
        w = dict(kind="single", grid=512, k_total=131072, horizon=50, envs=1, scaling="strong",
                 label=f"BASELINE configs[2]: 512x512 synthetic terrain, K=131072, T=50, samples sharded over {world} B200")
    elif name == "c3":
        if 64 % world:
            raise SystemExit("c3 shards 64 environments: --gpus must divide 64")
        w = dict(kind="batch", grid=64, k_total=4096, horizon=30, envs=64, scaling="strong",
                 label=f"BASELINE configs[3]: 64 environments x K=4096, T=30, own 64x64 map each, environments sharded "
                       f"over {world} B200 (no exchange)")
    elif name == "c4":
        if world != 1:
            raise SystemExit("c4 (stochastic slip) is a single-GPU configuration")
        w = dict(kind="stoch", grid=256, k_total=32768, horizon=50, envs=1, scaling="weak",
                 label="BASELINE configs[4]: stochastic slip (per-sample, per-lookup Normal draws), 256x256, K=32768, T=50")
    else:
        raise SystemExit(f"unknown config {name}")
    w["name"] = name
    return w


def config_of(w: dict, world: int) -> dict:
    """The `config` object of the JSON line -- the same in the native and the reference arm."""
    return {"workload": w["label"], "name": w["name"], "grid": w["grid"], "num_samples_total": w["k_total"],
            "horizon": w["horizon"], "num_envs": w["envs"], "n_gpus": world,
            "contract": "full: in-engine noise kept, recorded states + weights written",
            "l2": "flushed between timed steps (256 MiB write); per-step CUDA-event pairs"}


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


def kernel_source_hash() -> str:
    """Identity of the rollout kernel's source (comments and white space ignored): the ncu traffic figure is only quoted
    for the kernel it was taken on."""
    import re

    h = hashlib.sha256()
    for name in ("mppi_kernels.cuh", "mppi_math.cuh", "ptx_sm100.cuh"):
        with open(os.path.join(ROOT, "benchnav_b200", "csrc", name), "r", encoding="utf-8") as f:
            text = f.read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)   # block comments
        text = re.sub(r"//[^\n]*", "", text)                 # line comments
        h.update(re.sub(r"\s+", "", text).encode())
    return h.hexdigest()[:16]


def ncu_traffic(workload: str):
    """(per-launch DRAM bytes of the rollout kernel from the committed ncu capture, note).  Refused -- null -- when the
    capture was taken on a different version of the kernel source (profiles/rollout_traffic.json records the hash)."""
    try:
        with open(os.path.join(ROOT, "profiles", "rollout_traffic.json")) as f:
            d = json.load(f)
    except (OSError, ValueError):
        return None, "no ncu capture committed"
    entry = d.get(workload)
    if not entry:
        return None, f"no ncu capture for {workload}"
    if entry.get("kernel_source_hash") != kernel_source_hash():
        return None, (f"stale: capture taken on kernel source {entry.get('kernel_source_hash')}, "
                      f"current {kernel_source_hash()}")
    return entry.get("dram_bytes_per_launch"), f"ncu --set full, {entry.get('source', 'profiles/')}"


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


def build_problem(grid: int, seed: int = 0):
    import torch  # noqa: F401

    from benchnav_b200.synthetic import benchmark_problem

    return benchmark_problem(grid, RESOLUTION, seed=seed)


def stoch_problem(grid: int):
    """c4: the synthetic terrain's slip prediction Normal(mean, std); same start / goal / threshold as c1."""
    from benchnav_b200.synthetic import make_terrain

    terr = make_terrain(grid, RESOLUTION, 0)
    _, start, goal, thr = build_problem(grid)
    return terr["slip_mean"], terr["slip_std"], start, goal, thr


# ------------------------------------------------------------------------------------------ reference arm
def _bounded(est_s: float, steps: int, warmup: int, budget_s: float, k: int, horizon: int):
    """Shrink (K, T) of a CPU step so that steps + warmup of them fit the budget; the rate is scaled back linearly by
    the sampled fraction of the K x T rollout steps (the reference loop is a Python loop over T of [K]-wide ATen ops)."""
    frac = budget_s / max(est_s * (steps + warmup), 1e-9)
    k_run, t_run = k, horizon
    if frac < 1.0:
        t_run = max(1, int(horizon * frac))
        if horizon * frac < 1.0:
            k_run = max(1024, int(k * frac * horizon))
    return k_run, t_run, (k_run * t_run) / (k * horizon)


def cpu_port_rate(w: dict, steps: int, warmup: int, budget_s: float = 100.0):
    """The oracle port of the reference loop on the host cores for workload `w` (whole job: all samples / all
    environments of all ranks).  Returns (control iterations/s of the job, ms per job step, cores, sample text)."""
    import torch

    from oracle import mppi_oracle as orc

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g, k, t_h = w["grid"], w["k_total"], w["horizon"]
    sig = torch.tensor(SIGMAS)
    if w["kind"] == "batch":
        # configs[3]'s reference form: E separate solvers stepped in a Python loop; sample = the first few
        envs = w["envs"]
        solvers, starts = [], []
        for e in range(envs):
            risk, start, goal, thr = build_problem(g, seed=e)
            solvers.append(orc.OracleSolver(orc.make_problem(risk, RESOLUTION, goal.tolist(), thr), t_h, k, SIGMAS,
                                            LAMBDA, seed=SEED + e))
            starts.append(start)
        solvers[0].forward(starts[0])
        t0 = time.perf_counter()
        solvers[0].forward(starts[0])
        est = time.perf_counter() - t0
        e_run = int(max(1, min(envs, budget_s / max(est * (steps + warmup), 1e-9))))
        for _ in range(warmup):
            for e in range(e_run):
                solvers[e].forward(starts[e])
        t0 = time.perf_counter()
        for _ in range(steps):
            for e in range(e_run):
                solvers[e].forward(starts[e])
        dt = (time.perf_counter() - t0) / steps * (envs / e_run)
        sample = (f"{steps} steps of {e_run} of the {envs} environments (K={k}, T={t_h}, G={g} each; one port solver per "
                  f"environment in a Python loop) after {warmup} warm-up; time scaled by {envs}/{e_run}")
        return 1.0 / dt, dt * 1e3, cores, sample
    if w["kind"] == "stoch":
        mean, std, start, goal, thr = stoch_problem(g)
        p = orc.make_problem(mean, RESOLUTION, goal.tolist(), thr)
        p.slip_std = std
        gen = torch.Generator().manual_seed(SEED)

        def step(k_run, t_run, u_prev):
            noise = torch.randn(k_run, t_run, 2, generator=gen) * sig
            xi = torch.randn(k_run, 2 * t_run + 1, generator=gen)
            xi_opt = torch.randn(t_run, generator=gen)
            return orc.mppi_iteration(p, start, u_prev, noise, sig, LAMBDA, xi=xi, xi_opt=xi_opt)["u_opt"]
    else:
        risk, start, goal, thr = build_problem(g)
        p = orc.make_problem(risk, RESOLUTION, goal.tolist(), thr)
        gen = torch.Generator().manual_seed(SEED)

        def step(k_run, t_run, u_prev):
            noise = torch.randn(k_run, t_run, 2, generator=gen) * sig
            return orc.mppi_iteration(p, start, u_prev, noise, sig, LAMBDA)["u_opt"]

    # probe on a small slice to size the bounded sample (a full c2 step takes seconds)
    k_probe = min(k, 16384)
    step(k_probe, t_h, torch.zeros(t_h, 2))
    t0 = time.perf_counter()
    step(k_probe, t_h, torch.zeros(t_h, 2))
    est = (time.perf_counter() - t0) * (k / k_probe)
    k_run, t_run, scale = _bounded(est, steps, warmup, budget_s, k, t_h)
    u_prev = torch.zeros(t_run, 2)
    for _ in range(warmup):
        u_prev = step(k_run, t_run, u_prev)
    t0 = time.perf_counter()
    for _ in range(steps):
        u_prev = step(k_run, t_run, u_prev)
    dt = (time.perf_counter() - t0) / steps / scale
    sample = f"{steps} forward() calls of K={k_run}, T={t_run} on G={g} after {warmup} warm-up"
    sample += " (full workload)" if scale == 1.0 else f"; rate scaled by {k_run}*{t_run}/({k}*{t_h}) to the full workload"
    return 1.0 / dt, dt * 1e3, cores, sample


def reference_class_rate(w: dict, steps: int, warmup: int, budget_s: float = 60.0):
    """The UNMODIFIED reference `MPPI` class on the host cores, when the reference tree is importable (the build
    container; it does not travel to the GPU box).  Deterministic single-solver workloads only.  None otherwise."""
    ref_root = os.environ.get("BENCHNAV_REFERENCE", "/root/reference")
    if w["kind"] != "single" or not os.path.isdir(os.path.join(ref_root, "src")):
        return None
    import types

    import torch
    from torch.distributions import Normal

    sys.modules.setdefault("opensimplex", types.SimpleNamespace(seed=lambda s: None))
    for p in (os.path.join(ref_root, "src"), ref_root):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        from src.environments.grid_map import GridMap
        from src.planners.local_planners.mppi import MPPI as RefMPPI
        from src.simulator.problem_formulation.objectives import Objectives
        from src.simulator.problem_formulation.robot_model import UnicycleModel
        from src.simulator.problem_formulation.utils import ModelConfig
    except Exception:  # noqa: BLE001
        return None
    g, k, t_h = w["grid"], w["k_total"], w["horizon"]
    risk, start, goal, thr = build_problem(g)
    dists = {"predictions": Normal(risk, torch.full_like(risk, 0.1)), "latent_models": Normal(risk, torch.full_like(risk, 0.1))}
    gm = GridMap(g, RESOLUTION, tensors={"heights": torch.zeros(g, g)}, distributions=dists, instance_name="bench",
                 device="cpu")
    dyn = UnicycleModel(gm, ModelConfig("inference", "expected_value"), device="cpu")
    obj = Objectives(dyn, goal_pos=goal, stuck_threshold=thr)
    torch.set_num_threads(os.cpu_count() or 1)
    k_probe = min(k, 16384)
    probe = RefMPPI(t_h, k_probe, 3, 2, dyn, obj, torch.tensor(SIGMAS), LAMBDA, device=torch.device("cpu"), seed=SEED)
    with torch.no_grad():
        probe.forward(state=start)
        t0 = time.perf_counter()
        probe.forward(state=start)
    est = (time.perf_counter() - t0) * (k / k_probe)
    k_run, t_run, scale = _bounded(est, steps, warmup, budget_s, k, t_h)
    solver = RefMPPI(t_run, k_run, 3, 2, dyn, obj, torch.tensor(SIGMAS), LAMBDA, device=torch.device("cpu"), seed=SEED)
    with torch.no_grad():
        for _ in range(warmup):
            solver.forward(state=start)
        t0 = time.perf_counter()
        for _ in range(steps):
            solver.forward(state=start)
    dt = (time.perf_counter() - t0) / steps / scale
    return {"value": 1.0 / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
            "sample": f"{steps} forward() calls of the reference MPPI class, K={k_run}, T={t_run}, G={g}" +
                      ("" if scale == 1.0 else f"; rate scaled by {scale:.4f} to the full workload")}


def unit_factor(w: dict, world: int) -> int:
    """value = control iterations/s x this factor (weak scaling counts shard iterations, see the module docstring)."""
    return world if w["name"] == "c2w" else 1


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = resolve_workload(args.config, args.gpus)
    rate, ms, cores, sample = cpu_port_rate(w, args.steps, max(args.warmup, 1))
    value = rate * unit_factor(w, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_of(w, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "detail": {"arm": "oracle port of the reference PyTorch loop on the host cores (the reference is pure Python and "
                          "its packaging installs only src/simulator, see DESIGN.md: it cannot travel to the GPU box)",
                   "control_iters_per_sec": rate},
    }
    ref_cls = reference_class_rate(w, min(args.steps, 5), 1)
    if ref_cls is not None:
        line["detail"]["reference_class"] = ref_cls  # the unmodified class, timed beside the port when importable
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ native arm
def make_native(w: dict, dev, group, exchange: str, world: int, rank: int):
    """Build the solver of workload `w` for this rank.  Returns a dict of closures: step() = device-resident call,
    e2e_step() = host-buffer call (or None), plus the algorithmic bytes and sample count of ONE launch on this rank."""
    import torch

    from benchnav_b200 import MPPI, BatchedMPPI
    from benchnav_b200.problem import GoalObjectives, GridSpec, SlipDistribution, UnicycleProblem

    g, t_h = w["grid"], w["horizon"]
    sig = torch.tensor(SIGMAS)
    if w["kind"] == "batch":
        from benchnav_b200.dist import shard_range

        lo, hi = shard_range(w["envs"], rank, world)
        dyns, objs, states = [], [], []
        for e in range(lo, hi):
            risk, start, goal, thr = build_problem(g, seed=e)
            d = UnicycleProblem(GridSpec(g, RESOLUTION), risk)
            dyns.append(d)
            objs.append(GoalObjectives(d, goal, thr))
            states.append(start)
        solver = BatchedMPPI(t_h, w["k_total"], dyns, objs, sig, LAMBDA, device=dev, seed=SEED + lo)
        e_l = hi - lo
        st_dev = torch.stack(states).to(dev)
        st_pin = torch.stack(states).pin_memory()
        u_pin = torch.empty(e_l, t_h, 2).pin_memory()
        o_pin = torch.empty(e_l, 1, t_h + 1, 3).pin_memory()

        def e2e_step():
            solver.forward_host(st_pin, out=(u_pin, o_pin))

        return {"solver": solver, "step": lambda: solver.forward(st_dev), "e2e_step": e2e_step,
                "h2d": e_l * 12, "d2h": e_l * 4 * (2 * t_h + 3 * (t_h + 1)),
                "e2e_how": "BatchedMPPI.forward_host(states [E,3] in host memory, out=caller buffers): one staged H2D copy "
                           "of the states, the iteration -- every environment's last CTA stores u* [T,2] and the optimal "
                           "states [T+1,3] straight into pinned, device-mapped host memory -- and one stream "
                           "synchronisation, all inside the library; wall clock per step",
                "alg_bytes": e_l * algorithmic_bytes(w["k_total"], t_h, g), "units": e_l, "local_samples": w["k_total"],
                "parallelism": f"environment-shard x{world}: {e_l} environments per GPU, one launch, no exchange"}
    if w["kind"] == "stoch":
        mean, std, start, goal, thr = stoch_problem(g)
        dyn = UnicycleProblem(GridSpec(g, RESOLUTION, distributions={"predictions": SlipDistribution(mean, std)}), mean)
        solver = MPPI(t_h, w["k_total"], 3, 2, dyn, GoalObjectives(dyn, goal, thr), sig, LAMBDA, device=dev, seed=SEED,
                      stochastic_slip=True)
        channels = 2
    else:
        risk, start, goal, thr = build_problem(g)
        dyn = UnicycleProblem(GridSpec(g, RESOLUTION), risk)
        solver = MPPI(t_h, w["k_total"], 3, 2, dyn, GoalObjectives(dyn, goal, thr), sig, LAMBDA, device=dev, seed=SEED,
                      process_group=group, exchange=exchange)
        channels = 1
    st_dev = start.to(dev)
    st_pin = start.clone().pin_memory()
    outs = (torch.empty(t_h, 2).pin_memory(), torch.empty(1, t_h + 1, 3).pin_memory())
    k_l = solver._local_samples
    par = f"sample-shard x{world}" + ("" if world == 1 else (", fused exchange of the softmax partial inside the rollout "
                                                             "kernel over NVLink peer memory" if solver._fused_exchange
                                                             else ", NCCL all-gather + finalize kernel"))
    return {"solver": solver, "step": lambda: solver.forward(st_dev),
            "e2e_step": (lambda: solver.forward_host(st_pin, out=outs)) if (world == 1 or rank == 0) else
                        ((lambda: solver.forward_follow()) if solver._fused_exchange else None),
            "two_stage": ((lambda: solver.forward_action(st_pin, out=outs[0])), (lambda: solver.wait_states(outs[1])))
                         if world == 1 else None,
            "h2d": 12, "d2h": 4 * (2 * t_h + 3 * (t_h + 1)),
            "e2e_how": "forward_host(state, out=caller buffers): the 12-byte state rides in the launch packet (or a "
                       "pinned, device-mapped mailbox when pre-launched), the kernel stores u* and the optimal state "
                       "sequence into pinned mapped host memory and raises a completion word the host polls; wall clock" +
                       ("" if world == 1 else "; sharded: the control loop runs on rank 0 (its kernel broadcasts the state "
                        "to the other ranks over NVLink), every other rank calls forward_follow(), whose kernel waits on "
                        "the device for that state; rank 0's wall clock per step, its host buffers receive the results"),
            "alg_bytes": algorithmic_bytes(k_l, t_h, g, channels), "units": 1, "local_samples": k_l,
            "parallelism": par, "start": start, "state_dev": st_dev}


def run_native(args) -> None:
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run "
                         f"--nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    w = resolve_workload(args.config, world)
    sharded = w["kind"] == "single" and world > 1
    nat = make_native(w, dev, group if w["kind"] == "single" else None, args.exchange, world, rank)
    solver, step = nat["solver"], nat["step"]
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    steps, warmup = args.steps, max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_pass(fn, n: int, sync=barrier) -> float:
        """n calls, each preceded by an L2 flush, each bracketed by its own CUDA-event pair; returns milliseconds."""
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        for e in ev0 + ev1:  # torch creates the CUDA event at its first record(): do that here, not inside the first
            e.record()       # timed step, where the host then falls behind the device and the launch gap is timed
        sync()
        for i in range(n):
            flush.fill_(i & 0xFF)
            ev0[i].record()
            fn()
            ev1[i].record()
        sync()
        per_step = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
        if os.environ.get("BNV_BENCH_DUMP_STEPS"):  # debug aid: the per-step device times of this pass, microseconds
            print("per-step us:", " ".join(f"{t * 1e3:.1f}" for t in per_step[:64]), file=sys.stderr)
        return sum(per_step)

    # ---- parity of the sharded solver, on its very first step (driver-run evidence in every multi-GPU line)
    parity = None
    single_ref = None
    if sharded:
        from benchnav_b200 import MPPI
        from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem

        u0, _ = step()  # iteration 0 of the sharded solver
        torch.cuda.synchronize(dev)
        gathered = [torch.empty_like(u0) for _ in range(world)]
        dist.all_gather(gathered, u0)
        same = all(torch.equal(gathered[0], g_) for g_ in gathered[1:])
        if rank == 0:
            risk, start, goal, thr = build_problem(w["grid"])
            dyn = UnicycleProblem(GridSpec(w["grid"], RESOLUTION), risk)
            single_ref = MPPI(w["horizon"], w["k_total"], 3, 2, dyn, GoalObjectives(dyn, goal, thr),
                              torch.tensor(SIGMAS), LAMBDA, device=dev, seed=SEED)
            u_ref, _ = single_ref.forward(nat["state_dev"])  # iteration 0 of the unsharded solver: same Philox stream
            torch.cuda.synchronize(dev)
            diff = float((u_ref - u0).abs().max())
            ok = same and diff <= 5e-6
            parity = {"parity_check": "ok" if ok else "FAILED", "ranks_bit_equal": bool(same),
                      "max_abs_du_vs_unsharded": diff, "tolerance": 5e-6,
                      "what": "first step: u* all-gathered over ranks is bit-identical, and equals the unsharded "
                              "single-GPU solver's u* (same seed, same Philox stream by global sample index)"}

    for _ in range(warmup):
        flush.fill_(1)
        step()
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = solver.launch_count
    total_ms = max_over_ranks(timed_pass(step, steps))
    launches = solver.launch_count - launches0
    clocks = sampler.stop() if sampler else None

    # ---- second pass: rollout kernel alone (engine-side event pairs) -> roofline
    n_k = min(steps, 4096)
    solver.kernel_timing(n_k)
    timed_pass(step, n_k)
    kern_ms, kern_n = solver.kernel_time()
    solver.kernel_timing(0)
    kern_s = kern_ms / max(kern_n, 1) * 1e-3

    # ---- back-to-back (no flush): how far launch gaps / cold L2 matter
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    hot_ms = max_over_ranks(e0.elapsed_time(e1) / steps)

    # ---- end to end through the host-buffer call
    e2e = None
    if world > 1 and w["kind"] == "single":  # every rank must take part (or none): agree on it
        ok = torch.tensor([1 if nat["e2e_step"] is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            nat["e2e_step"] = None
    if nat["e2e_step"] is not None:
        n_e = min(steps, 2000)
        cur = torch.cuda.current_stream(dev)
        e2e_step = nat["e2e_step"]

        def e2e_pass(n: int) -> float:
            """n steps: L2 flush (not timed; only the flush's own stream is synchronised -- a pre-launched kernel is
            meant to be waiting on the device at that point), then the timed host-buffer call."""
            for _ in range(3):
                e2e_step()
            total = 0.0
            follower = sharded and rank != 0  # enqueues only: its kernels wait on the device for the leader's state
            for i in range(n):
                flush.fill_(i & 0xFF)
                if not follower:
                    cur.synchronize()
                t0 = time.perf_counter()
                e2e_step()
                total += time.perf_counter() - t0
            return total

        barrier()
        acc_plain = e2e_pass(n_e)
        if sharded:  # followers only enqueue (they run ahead of the leader on purpose): the leader's clock counts
            barrier()
            solver.check()
            acc_plain = max_over_ranks(acc_plain if rank == 0 else 0.0)
        else:
            acc_plain = max_over_ranks(acc_plain)
        factor = unit_factor(w, world)
        e2e = {"value": n_e / acc_plain * factor, "unit": UNIT, "h2d_bytes_per_step": nat["h2d"],
               "d2h_bytes_per_step": nat["d2h"], "steps": n_e, "mode": "plain launches", "timing": nat["e2e_how"]}
        if w["kind"] != "batch" and world == 1:
            # the same call with the next iteration's kernel pre-launched (opt-in, solver.prelaunch()): the kernel is
            # resident and polling a host-mapped mailbox when the state arrives.  Reported beside the plain number; the
            # headline `value` of e2e is the better of the two and `mode` says which.
            try:
                solver.prelaunch(True)
                acc_pre = e2e_pass(n_e)
                e2e["value_prelaunched"] = n_e / acc_pre
                e2e["value_plain_launch"] = n_e / acc_plain
                if acc_pre < acc_plain:
                    e2e["value"], e2e["mode"] = n_e / acc_pre, "pre-launched iterations (solver.prelaunch())"
                if nat.get("two_stage"):
                    # two-stage form of the same call: forward_action returns on the kernel's first completion word (u*
                    # in host memory, what the loop needs to step the environment), wait_states on the second (the
                    # optimal state sequence).  Both halves are timed; the headline stays the full result.
                    act, rest = nat["two_stage"]
                    for _ in range(3):
                        act()
                        rest()
                    t_act = t_rest = 0.0
                    for i in range(n_e):
                        flush.fill_(i & 0xFF)
                        cur.synchronize()
                        t0 = time.perf_counter()
                        act()
                        t1 = time.perf_counter()
                        rest()
                        t_rest += time.perf_counter() - t1
                        t_act += t1 - t0
                    e2e["two_stage"] = {"action_us": t_act / n_e * 1e6, "states_after_action_us": t_rest / n_e * 1e6,
                                        "d2h_bytes_action": 4 * 2 * w["horizon"],
                                        "what": "pre-launched; forward_action(state) -> u* [T,2] in host memory, then "
                                                "wait_states() -> optimal states [1,T+1,3]; mean wall clock of each half"}
            except Exception as exc:  # noqa: BLE001 -- keep the bench line
                e2e["prelaunch_error"] = str(exc)
            finally:
                try:
                    solver.prelaunch(False)
                except Exception:  # noqa: BLE001
                    pass
            # the reference-style call: forward(host state) returning CUDA tensors, then .cpu() on both results --
            # what Tutorial 3.3's loop does (test/test_mppi.py:176-181)
            n_r = min(n_e, 500)
            acc_ref_style = 0.0
            start_host = nat["start"]
            for i in range(n_r + 3):
                flush.fill_(i & 0xFF)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                u_, o_ = solver.forward(start_host)
                u_.cpu()
                o_.cpu()
                if i >= 3:
                    acc_ref_style += time.perf_counter() - t0
            e2e["value_forward_then_cpu"] = n_r / acc_ref_style

    # ---- strong-scaling datum: ONE GPU on the same total workload, same method (rank 0 only; the others wait)
    single_same = None
    if sharded:
        barrier()
        if rank == 0:
            n_s = min(steps, 500)
            st = nat["state_dev"]
            for _ in range(3):
                single_ref.forward(st)
            ms_1 = timed_pass(lambda: single_ref.forward(st), n_s, sync=lambda: torch.cuda.synchronize(dev)) / n_s
            single_same = {"single_gpu_same_total_ms": ms_1, "single_gpu_steps": n_s,
                           "single_gpu_launch": single_ref.launch_geometry}
        barrier()

    # ---- N = 1 of the weak-scaling family: the SAME per-GPU workload the N > 1 lines run (K = 16384 on the 512x512
    # terrain), so that scaling efficiency has one baseline (the headline c1 is the 256x256 configuration)
    weak_base = None
    if world == 1 and w["name"] == "c1":
        from benchnav_b200 import MPPI
        from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem

        wb = resolve_workload("c2w", 1)
        risk_b, start_b, goal_b, thr_b = build_problem(wb["grid"])
        dyn_b = UnicycleProblem(GridSpec(wb["grid"], RESOLUTION), risk_b)
        sol_b = MPPI(wb["horizon"], wb["k_total"], 3, 2, dyn_b, GoalObjectives(dyn_b, goal_b, thr_b), torch.tensor(SIGMAS),
                     LAMBDA, device=dev, seed=SEED)
        st_b = start_b.to(dev)
        for _ in range(3):
            sol_b.forward(st_b)
        n_b = min(steps, 1000)
        ms_b = timed_pass(lambda: sol_b.forward(st_b), n_b) / n_b
        weak_base = {"workload": wb["label"], "ms_per_step": ms_b, "iters_per_sec": 1e3 / ms_b, "steps": n_b}
        sol_b.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / steps
    control_rate = 1e3 / ms_per_step
    factor = unit_factor(w, world)
    peak, peak_src = hbm_peak()
    achieved = nat["alg_bytes"] / kern_s / 1e9 if kern_s > 0 else 0.0
    traffic, traffic_note = ncu_traffic(f"{w['name']}_G{w['grid']}_K{nat['local_samples']}_T{w['horizon']}")
    detail = {"parallelism": nat["parallelism"], "control_iters_per_sec": control_rate,
              "back_to_back_ms_per_step": hot_ms,
              "rollout_steps_per_sec": control_rate * w["k_total"] * w["horizon"] * w["envs"],
              "value_is": "control iterations/s" + (f" x {world} (16384-sample shard iterations/s, weak scaling)" if factor > 1 else "")}
    if hasattr(solver, "launch_geometry"):
        detail["launch"] = solver.launch_geometry
    if w["kind"] == "batch":
        detail["env_iters_per_sec"] = control_rate * w["envs"]
    if weak_base is not None:
        detail["weak_scaling_n1"] = weak_base
    if single_same is not None:
        detail.update(single_same)
        detail["speedup_vs_1gpu"] = single_same["single_gpu_same_total_ms"] / ms_per_step
    line = {
        "metric": METRIC, "value": control_rate * factor, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(w, world),
        "roofline": {"bound": "hbm", "kernel": "bnv::rollout_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                     "algorithmic_bytes": nat["alg_bytes"], "kernel_us": kern_s * 1e6, "launches_timed": kern_n,
                     "peak_source": peak_src, "kernel_source_hash": kernel_source_hash()},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "detail": detail,
    }
    if parity is not None:
        line.update({"parity_check": parity["parity_check"]})
        detail["parity"] = parity
    if e2e is not None:
        line["e2e"] = e2e
    else:
        line["e2e"] = {"value": control_rate * factor, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                       "note": "sample-sharded multi-GPU: device-resident state; no host-buffer number claimed"}
    if world == 1:
        rate, ms, cores, sample = cpu_port_rate(w, 5, 2, budget_s=25.0)
        line["cpu_baseline"] = {"value": rate * factor, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", choices=("native", "reference"), default="native")
    ap.add_argument("--config", choices=("auto", "c1", "c2w", "c2", "c3", "c4"), default="auto",
                    help="workload (see the module docstring); auto = c1 on one GPU, c2w on several")
    ap.add_argument("--exchange", choices=("p2p", "nccl"), default="p2p",
                    help="multi-GPU softmax exchange: fused over NVLink peer memory (default) or NCCL all-gather")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
