#!/bin/bash
python - <<'PY'
import sys, time, torch
sys.path.insert(0,'examples'); sys.path.insert(0,'.')
import closed_loop as m
from benchnav_b200 import BatchedMPPI, BatchedPlanetaryEnv
dev=torch.device('cuda')
def trial(fused, external, envs=8, samples=4096, horizon=30, n=1000):
    dyns, objs, gms, start, goal = m.build(envs)
    planner = BatchedMPPI(horizon, samples, dyns, objs, torch.tensor([0.5, 0.5]), 0.5, device=dev, seed=0)
    env = BatchedPlanetaryEnv(gms, start, goal, delta_t=0.1, time_limit=100, stuck_threshold=0.1, goal_threshold=1.0, seed=0, device=dev, graph_capturable=True)
    state = env.reset(seed=0)
    done = torch.zeros(envs, dtype=torch.uint8, device=dev); stg = torch.full((envs,), -1, dtype=torch.long, device=dev)
    step_no = torch.zeros((), dtype=torch.long, device=dev); coll = torch.zeros(envs, horizon+1, dtype=torch.uint8, device=dev)
    zero = torch.zeros(envs, 2, device=dev)
    def cs():
        a, s = planner.forward(state)
        if fused:
            env.closed_loop_step(a, s, done, stg, step_no, coll, planner=planner if external else None)
        else:
            a0 = torch.where(done.bool().unsqueeze(1), zero, a[:, 0, :]); _, _, term, _ = env.step(a0)
            coll.copy_(env.collision_check(s[:, 0])); step_no.add_(1)
            stg.copy_(torch.where(term & ~done.bool(), step_no, stg)); done.copy_(done.bool() | term)
    planner.graph_capturable(True, external_advance=external)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        cs(); g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side): cs()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    for _ in range(20): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): g.replay()
    t_enq = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
    print(f"fused={fused} external={external}: enqueue {t_enq/n*1e6:.1f} us/replay, wall {wall/n*1e6:.1f} us/step, events {e0.elapsed_time(e1)/n*1e3:.1f} us/step, launches/forward {planner.launch_count}")
trial(True, True); trial(True, False); trial(False, False)
PY
