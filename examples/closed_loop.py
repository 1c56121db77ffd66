#!/usr/bin/env python
"""Tutorial 3.3's control loop on the device: E planetary environments and their MPPI planners advance together,
one planner launch + one environment launch per control step, robot states never leaving HBM.

    python examples/closed_loop.py [--envs 8] [--samples 4096] [--horizon 30] [--max-steps 600]

Mirrors the loop of the reference notebook (solver.forward -> env.step -> env.collision_check ->
solver.get_top_samples; notebooks/tutorial_3_3, test/test_mppi.py:171-198) with `BatchedMPPI` and
`BatchedPlanetaryEnv` standing in for `MPPI` and `PlanetaryEnv`; rendering is out of scope.
"""

from __future__ import annotations

import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from benchnav_b200 import BatchedMPPI, BatchedPlanetaryEnv  # noqa: E402
from benchnav_b200.problem import GoalObjectives, GridSpec, SlipDistribution, UnicycleProblem  # noqa: E402
from benchnav_b200.synthetic import make_terrain  # noqa: E402


def build(envs: int, grid: int = 128, resolution: float = 0.5, slip_scale: float = 0.4):
    """E synthetic terrains (different seeds); the planner sees the predicted slip (mean as risk map), the
    environment draws from the latent slip model."""
    dyns, objs, gms = [], [], []
    lim = grid * resolution
    start = torch.tensor([[0.15 * lim, 0.15 * lim]] * envs)
    goal = torch.tensor([[0.4 * lim, 0.45 * lim]] * envs)
    for e in range(envs):
        terr = make_terrain(grid, resolution, seed=100 + e)
        mean = (terr["slip_mean"] * slip_scale).clamp(0.0, 0.95)
        std = terr["slip_std"] * 0.5
        for pos in (start[e], goal[e]):  # keep the start and goal cells drivable, as PlanetaryEnv requires
            cx, cy = int(pos[0] / resolution), int(pos[1] / resolution)
            mean[cy - 2:cy + 3, cx - 2:cx + 3] = mean[cy - 2:cy + 3, cx - 2:cx + 3].clamp(max=0.2)
        d = SlipDistribution(mean, std)
        gm = GridSpec(grid, resolution, distributions={"predictions": d, "latent_models": d})
        dyn = UnicycleProblem(gm, mean)
        dyns.append(dyn)
        objs.append(GoalObjectives(dyn, goal[e], 0.3))
        gms.append(gm)
    return dyns, objs, gms, start, goal


def run(envs: int = 8, samples: int = 4096, horizon: int = 30, max_steps: int = 900, seed: int = 0, verbose: bool = True,
        use_graph: bool = True, fused: bool = True):
    """Drive every environment to its goal.  `use_graph`: capture ONE control step in a CUDA graph and replay it -- one
    graph launch per control step.  `fused`: everything between two planner calls (environment step, collision check,
    the loop's books) is ONE kernel (`BatchedPlanetaryEnv.closed_loop_step`), so a control step is two kernels; otherwise
    the same sequence through the separate calls (~15 small launches), bit-identical."""
    dev = torch.device("cuda")
    dyns, objs, gms, start, goal = build(envs)
    planner = BatchedMPPI(horizon, samples, dyns, objs, torch.tensor([0.5, 0.5]), 0.5, device=dev, seed=seed)
    env = BatchedPlanetaryEnv(gms, start, goal, delta_t=0.1, time_limit=100, stuck_threshold=0.1, goal_threshold=1.0,
                              seed=seed, device=dev, graph_capturable=use_graph or fused)
    state = env.reset(seed=seed)  # updated in place by env.step: the graph's static input
    done = torch.zeros(envs, dtype=torch.uint8, device=dev)
    steps_to_goal = torch.full((envs,), -1, dtype=torch.long, device=dev)
    step_no = torch.zeros((), dtype=torch.long, device=dev)
    zero = torch.zeros(envs, 2, device=dev)
    collisions = torch.zeros(envs, horizon + 1, dtype=torch.uint8, device=dev)
    external = fused and use_graph  # the fused kernel also advances the planner's iteration counter

    def control_step():
        actions, state_seqs = planner.forward(state)                       # [E,T,2], [E,1,T+1,3]
        if fused:
            env.closed_loop_step(actions, state_seqs, done, steps_to_goal, step_no, collisions,
                                 planner=planner if external else None)
            return
        a0 = torch.where(done.bool().unsqueeze(1), zero, actions[:, 0, :])  # arrived robots stop
        _, _, terminated, _ = env.step(a0)
        collisions.copy_(env.collision_check(state_seqs[:, 0]))            # [E,T+1] on the planned trajectory
        step_no.add_(1)
        steps_to_goal.copy_(torch.where(terminated & ~done.bool(), step_no, steps_to_goal))  # no host sync in the loop
        done.copy_(done.bool() | terminated)

    graph = None
    if use_graph:
        planner.graph_capturable(True, external_advance=external)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            control_step()  # warm-up outside the capture (allocator, lazy module loads)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                control_step()
        torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t_steady, n_steady0 = t0, 0
    n_steps = 0
    for step in range(max_steps):
        if graph is not None:
            graph.replay()
        else:
            control_step()
        n_steps += 1
        if step % 50 == 49:  # one host sync every 50 steps
            all_done = bool(done.bool().all())
            if step == 49:  # the first replays carry the graph's upload: the steady-state rate is taken after them
                t_steady, n_steady0 = time.perf_counter(), n_steps
            if all_done:
                break
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    per_step = (time.perf_counter() - t_steady) / max(n_steps - n_steady0, 1) if n_steps > n_steady0 else wall / n_steps
    top_states, top_weights = planner.get_top_samples(min(100, samples))
    dist = (state[:, :2] - goal.to(dev)).norm(dim=1)
    if verbose:
        print(f"{envs} environments, K={samples}, T={horizon}, {'CUDA graph' if graph is not None else 'host-issued launches'}, "
              f"{'fused loop kernel' if fused else 'separate calls'}: "
              f"{n_steps} control steps in {wall * 1e3:.1f} ms; steady state {per_step * 1e6:.1f} us per step of all environments")
        print("steps to goal per environment:", steps_to_goal.tolist())
        print("final distance to goal [m]:", [round(float(x), 2) for x in dist])
        print("planned-trajectory collisions flagged at the last step:", int(collisions.sum()))
    return steps_to_goal.cpu(), dist.cpu(), per_step


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8)
    ap.add_argument("--samples", type=int, default=4096)
    ap.add_argument("--horizon", type=int, default=30)
    ap.add_argument("--max-steps", type=int, default=900)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--unfused", action="store_true")
    a = ap.parse_args()
    run(a.envs, a.samples, a.horizon, a.max_steps, use_graph=not a.no_graph, fused=not a.unfused)
