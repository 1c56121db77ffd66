#!/usr/bin/env python
"""Device-timed rates of the widened path on one B200 (not the bench.py headline): BASELINE config 3's per-GPU share
(8 environments x K=4096 x T=30) and its full 64-environment form, config 4 (stochastic slip, 256x256, K=32768, T=50),
and the rows either side of the iteration (risk map, environment step, collision check, DWA).
Per-step CUDA-event pairs, 256 MiB L2 flush between steps.  One JSON object per line."""

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from benchnav_b200 import DWA, MPPI, BatchedMPPI, BatchedPlanetaryEnv, infer_risk_map  # noqa: E402
from benchnav_b200.problem import GoalObjectives, GridSpec, SlipDistribution, UnicycleProblem  # noqa: E402
from benchnav_b200.synthetic import benchmark_problem, make_terrain  # noqa: E402

DEV = torch.device("cuda")
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, n=300, warm=20):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for a, b in ev:
        FLUSH.fill_(1)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return sum(ts) / n, ts[n // 2]


def batch_case(E, K, T, g=64):
    dyns, objs, states = [], [], []
    for e in range(E):
        risk, start, goal, thr = benchmark_problem(g, 0.5, seed=e)
        d = UnicycleProblem(GridSpec(g, 0.5), risk)
        dyns.append(d)
        objs.append(GoalObjectives(d, goal, thr))
        states.append(start)
    solver = BatchedMPPI(T, K, dyns, objs, torch.tensor([0.5, 0.5]), 0.5, device=DEV, seed=1)
    st = torch.stack(states).to(DEV)
    mean_ms, med_ms = timed(lambda: solver.forward(st))
    byt = E * (8 * K * T + 12 * K * (T + 1) + 4 * K + 4 * g * g + 8 * T + 12 * (T + 1))
    return {"case": f"config3 batched: E={E} envs x K={K} x T={T}, G={g}", "ms_per_step": mean_ms, "median_ms": med_ms,
            "env_iters_per_sec": E / mean_ms * 1e3, "rollout_steps_per_sec": E * K * T / mean_ms * 1e3,
            "algorithmic_GBps": byt / mean_ms / 1e6, "launches_per_step": 1}


def stoch_case(K=32768, T=50, g=256):
    terr = make_terrain(g, 0.5, 0)
    mean, std = terr["slip_mean"], terr["slip_std"]
    d = SlipDistribution(mean, std)
    dyn = UnicycleProblem(GridSpec(g, 0.5, distributions={"predictions": d}), mean)
    obj = GoalObjectives(dyn, torch.tensor([0.375 * g * 0.5] * 2), 0.3)
    solver = MPPI(T, K, 3, 2, dyn, obj, torch.tensor([0.5, 0.5]), 0.5, device=DEV, seed=1, stochastic_slip=True)
    st = torch.tensor([8.0, 8.0, 0.785398], device=DEV)
    mean_ms, med_ms = timed(lambda: solver.forward(st))
    byt = 8 * K * T + 12 * K * (T + 1) + 4 * K + 8 * g * g + 8 * T + 12 * (T + 1)
    return {"case": f"config4 stochastic slip: G={g}, K={K}, T={T}", "ms_per_step": mean_ms, "median_ms": med_ms,
            "iters_per_sec": 1e3 / mean_ms, "rollout_steps_per_sec": K * T / mean_ms * 1e3,
            "algorithmic_GBps": byt / mean_ms / 1e6}


def single_case(K, T, g):
    risk, start, goal, thr = benchmark_problem(g, 0.5, seed=0)
    dyn = UnicycleProblem(GridSpec(g, 0.5), risk)
    solver = MPPI(T, K, 3, 2, dyn, GoalObjectives(dyn, goal, thr), torch.tensor([0.5, 0.5]), 0.5, device=DEV, seed=1)
    st = start.to(DEV)
    mean_ms, med_ms = timed(lambda: solver.forward(st))
    byt = 8 * K * T + 12 * K * (T + 1) + 4 * K + 4 * g * g + 8 * T + 12 * (T + 1)
    return {"case": f"single solver: G={g}, K={K}, T={T}", "ms_per_step": mean_ms, "median_ms": med_ms,
            "iters_per_sec": 1e3 / mean_ms, "algorithmic_GBps": byt / mean_ms / 1e6}


def aux_cases(g=256):
    out = []
    terr = make_terrain(g, 0.5, 0)
    mean, std = terr["slip_mean"].to(DEV), terr["slip_std"].to(DEV)
    for metric, method in (("cvar", "closed_form"), ("cvar", "monte_carlo"), ("var", "monte_carlo")):
        m, med = timed(lambda: infer_risk_map(mean, std, metric, 0.9, method=method, num_samples=1000), n=30, warm=3)
        out.append({"case": f"risk map {metric} {method}, G={g}, 1000 draws/cell", "ms": m, "median_ms": med})

    class GM:
        grid_size, resolution, x_limits, y_limits = g, 0.5, (0.0, g * 0.5), (0.0, g * 0.5)
        distributions = {"latent_models": SlipDistribution(mean, std)}

    E = 64
    env = BatchedPlanetaryEnv([GM] * E, torch.full((E, 2), 8.0), torch.full((E, 2), 48.0), device=DEV)
    acts = torch.rand(E, 2, device=DEV)
    m, med = timed(lambda: env.step(acts), n=200)
    out.append({"case": f"PlanetaryEnv.step x {E} envs (one launch)", "ms": m, "median_ms": med})
    top = torch.rand(E, 500 * 31, 3, device=DEV) * 100
    m, med = timed(lambda: env.collision_check(top), n=200)
    out.append({"case": f"collision_check of 500 top samples x 31 states x {E} envs", "ms": m, "median_ms": med})
    risk, start, goal, thr = benchmark_problem(g, 0.5, seed=0)
    dyn = UnicycleProblem(GridSpec(g, 0.5), risk)
    dwa = DWA(50, 3, 2, dyn, GoalObjectives(dyn, goal, thr), torch.tensor([0.5, 1.5]), 0.1, device=DEV)
    st = start.to(DEV)
    m, med = timed(lambda: dwa.forward(st), n=200)
    out.append({"case": "DWA.forward, 10x10 actions, T=50 (4 launches)", "ms": m, "median_ms": med})
    return out


if __name__ == "__main__":
    rows = [single_case(16384, 50, 256), single_case(4096, 30, 64), batch_case(8, 4096, 30), batch_case(64, 4096, 30),
            stoch_case(), stoch_case(K=16384), single_case(32768, 50, 256)] + aux_cases()
    for r in rows:
        print(json.dumps(r))
