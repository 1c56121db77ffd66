"""CPU-side checks of the widened boundary: argument validation of the new entry points happens before the device is
touched, the host mirrors refuse to run without CUDA (no CPU fallback), and the data carriers expose what the
introspection reads.  No compute is attempted."""

import ctypes as C

import pytest
import torch

from benchnav_b200 import _cabi
from benchnav_b200.problem import GoalObjectives, GridSpec, SlipDistribution, UnicycleProblem

cpu_only = pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-box behaviour")


def _grid(**kw):
    base = dict(grid_size=16, pitch=16, resolution=0.5, x_min=0.0, x_max=8.0, y_min=0.0, y_max=8.0)
    base.update(kw)
    return _cabi.Grid(**base)


def test_aux_entry_points_reject_bad_arguments():
    lib = _cabi.load()
    one = C.c_void_p(16)  # never dereferenced: validation fails first
    f2 = (C.c_float * 2)(0.0, 1.0)
    cases = [
        (lambda: lib.bnv_trav_lookup(None, one, None, 0, 0, one, 4, 3, None, 0, 0, None, 0.1, one, None, None), b"null grid"),
        (lambda: lib.bnv_trav_lookup(C.byref(_grid(pitch=8)), one, None, 0, 0, one, 4, 3, None, 0, 0, None, 0.1, one, None, None), b"pitch"),
        (lambda: lib.bnv_trav_lookup(C.byref(_grid()), one, None, 0, 0, one, 4, 1, None, 0, 0, None, 0.1, one, None, None), b"size"),
        (lambda: lib.bnv_trav_lookup(C.byref(_grid()), one, None, 0, 0, one, 4, 3, None, 0, 0, None, 0.1, None, None, None), b"null"),
        (lambda: lib.bnv_env_step(C.byref(_grid(resolution=0.0)), one, one, 0, 1, one, one, one, None, 0, 0, None, f2, f2, 0.1, 1.0, one, one, None), b"resolution"),
        (lambda: lib.bnv_env_step(C.byref(_grid()), one, one, 0, 0, one, one, one, None, 0, 0, None, f2, f2, 0.1, 1.0, one, one, None), b"size"),
        (lambda: lib.bnv_risk_map(3, 0.9, 0, one, one, 16, None, 100, 0, one, None, None), b"metric"),
        (lambda: lib.bnv_risk_map(1, 1.5, 0, one, one, 16, None, 100, 0, one, None, None), b"confidence"),
        (lambda: lib.bnv_risk_map(2, 0.9, 1, one, one, 16, None, 0, 0, one, None, None), b"num_samples"),
        (lambda: lib.bnv_risk_map(2, 0.9, 7, one, one, 16, None, 10, 0, one, None, None), b"method"),
        (lambda: lib.bnv_dwa_actions(one, f2, f2, f2, 0.1, 0, 10, 5, one, one, None), b"size"),
        (lambda: lib.bnv_mppi_set_problem_ex(None, one, None, 16, 16, 0, 0.5, 0.0, 8.0, 0.0, 8.0, f2, 0.3, None), b"null"),
        (lambda: lib.bnv_mppi_forward_ex(None, one, None, None, None, one, one, None), b"null"),
        (lambda: lib.bnv_mppi_draw_xi(None, 0, one, one, None), b"null"),
        (lambda: lib.bnv_mppi_argmin(None, one, one, one, None, None), b"null"),
        (lambda: lib.bnv_mppi_dwa_subgoal(None, one, 3, one, one, 1.0, one, None), b"null"),
    ]
    for call, needle in cases:
        assert call() == -1
        assert needle in lib.bnv_last_error(), lib.bnv_last_error()
    out = (C.c_int32 * 4)()
    assert lib.bnv_mppi_launch_geometry(None, out) == -1
    cfg = _cabi.MppiCfg(num_samples=64, horizon=8, lambda_=0.5, dt=0.1, world_size=2, rank=0, num_envs=4)
    for i in range(2):
        cfg.sigma[i], cfg.u_max[i] = 0.5, 1.0
    h = C.c_void_p()
    assert lib.bnv_mppi_create(C.byref(h), C.byref(cfg)) == -1 and b"world_size == 1" in lib.bnv_last_error()
    cfg.world_size, cfg.num_envs = 1, 70000
    assert lib.bnv_mppi_create(C.byref(h), C.byref(cfg)) == -1 and b"num_envs" in lib.bnv_last_error()


@cpu_only
def test_host_mirrors_fail_loudly_without_a_gpu():
    from benchnav_b200 import DWA, BatchedMPPI, BatchedPlanetaryEnv, infer_risk_map
    from benchnav_b200 import env as benv

    mean, std = torch.rand(16, 16), torch.rand(16, 16) * 0.1
    d = SlipDistribution(mean, std)
    grid = GridSpec(16, 0.5, distributions={"predictions": d, "latent_models": d})
    dyn = UnicycleProblem(grid, mean)
    obj = GoalObjectives(dyn, torch.tensor([4.0, 4.0]), 0.3)
    with pytest.raises(RuntimeError):
        BatchedMPPI(10, 64, [dyn, dyn], [obj, obj], torch.tensor([0.5, 0.5]), 0.5)
    with pytest.raises(RuntimeError):
        DWA(10, 3, 2, dyn, obj, torch.tensor([0.5, 1.5]), 0.1)
    with pytest.raises(RuntimeError):
        BatchedPlanetaryEnv([grid], torch.tensor([[2.0, 2.0]]), torch.tensor([[6.0, 6.0]]))
    with pytest.raises(RuntimeError):
        infer_risk_map(mean, std, "var", 0.9)
    with pytest.raises(RuntimeError):
        benv.collision_check(grid, torch.rand(3, 4, 3), 0.1, mean=mean, std=std)


def test_mirror_argument_checks_follow_the_reference():
    from benchnav_b200 import infer_risk_map
    from benchnav_b200.mppi import _slip_distribution

    mean, std = torch.rand(16, 16), torch.rand(16, 16)
    with pytest.raises(AssertionError):  # utils.py:18-24
        infer_risk_map(mean, std, "median")
    with pytest.raises(AssertionError):  # utils.py:27-31
        infer_risk_map(mean, std, "cvar", 1.2)
    d = SlipDistribution(mean, std)
    grid = GridSpec(16, 0.5, distributions={"predictions": d})
    m, s = _slip_distribution(grid)
    assert m is mean and s is std
    with pytest.raises(TypeError):
        _slip_distribution(GridSpec(16, 0.5))
    with pytest.raises(ValueError):
        _slip_distribution(GridSpec(8, 0.5, distributions={"predictions": d}))
