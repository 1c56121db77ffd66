"""CPU-side checks of the drop-in boundary: the library builds/loads without a GPU and exports every symbol
declared in include/bnv_mppi.h; argument validation that needs no device works; no compute is attempted."""

import ctypes as C
import os
import subprocess

import pytest
import torch

from benchnav_b200 import _cabi, build


def test_library_builds_and_loads():
    path = build.ensure_built()
    assert os.path.exists(path)
    lib = _cabi.load()
    assert lib.bnv_abi_version() == _cabi.ABI_VERSION


def test_every_declared_symbol_is_exported_and_bound():
    declared = _cabi.declared_symbols()
    assert len(declared) >= 20
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in bnv_mppi.h but not exported"
    assert set(declared) == set(_cabi._SIGNATURES), "ctypes binding and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", build.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(declared) <= exported


def test_cfg_struct_layout_matches_header():
    # int32 K, int32 T, float[2], float, float[2], float[2], float, (pad) uint64, int32 x3, uint32, int32, (pad)
    assert C.sizeof(_cabi.MppiCfg) == 72
    assert _cabi.MppiCfg.seed.offset == 40 and _cabi.MppiCfg.flags.offset == 60 and _cabi.MppiCfg.num_envs.offset == 64
    assert C.sizeof(_cabi.Grid) == 28


def test_sm100a_only_sass_with_tma_and_bulk_copies():
    """The shipped cubin is sm_100a and the rollout kernel really uses TMA (UTMALDG) and bulk copies (UBLKCP)."""
    sass = subprocess.run(["cuobjdump", "-sass", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    funcs = sass.split("Function : ")[1:]
    rollouts = [f for f in funcs if "rollout_kernel" in f.splitlines()[0]]
    # latency variant -- single solver: kPatch x kPow2 x kRecord x kFastAngles x kPhilox = 32; stochastic and/or batched
    # modes (record + fast angles only): 3 x kPatch x kPow2 x kPhilox = 24; wide variant (fast angles only) -- single
    # solver: kPatch x kPow2 x kRecord x kPhilox = 16; stochastic and/or batched: 24
    assert len(rollouts) == 96
    n_wide = 0
    for f in rollouts:
        name = f.splitlines()[0]
        patch = "rollout_kernelILb1E" in name
        wide = name.split("EEEv")[0].endswith("Lb1")  # last template flag
        n_wide += wide
        assert ("UTMALDG.2D" in f) == patch, name  # the TMA window load exists exactly in the kPatch variants
        # bulk slab copies in the latency variant; the wide variant flushes its chunks with coalesced vector stores
        assert ("UBLKCP" in f) == (not wide), name
        assert "HMMA" not in f and "UTCHMMA" not in f  # no tensor cores on this path
    assert n_wide == 40


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-box behaviour")
def test_fails_loudly_without_a_gpu():
    """No CPU fallback: creating a handle or a solver without CUDA raises, it does not degrade."""
    lib = _cabi.load()
    h = C.c_void_p()
    cfg = _cabi.MppiCfg(num_samples=64, horizon=8, lambda_=0.5, dt=0.1, world_size=1)
    for i in range(2):
        cfg.sigma[i], cfg.u_max[i] = 0.5, 1.0
    rc = lib.bnv_mppi_create(C.byref(h), C.byref(cfg))
    assert rc == -2 and lib.bnv_last_error()
    with pytest.raises(_cabi.BnvError):
        _cabi.check(rc)

    from benchnav_b200 import MPPI
    from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem

    dyn = UnicycleProblem(GridSpec(16, 0.5), torch.zeros(16, 16))
    obj = GoalObjectives(dyn, torch.tensor([4.0, 4.0]), 0.3)
    with pytest.raises(RuntimeError):
        MPPI(10, 64, 3, 2, dyn, obj, torch.tensor([0.5, 0.5]), 0.5)


def test_invalid_arguments_are_rejected_before_touching_the_device():
    lib = _cabi.load()
    h = C.c_void_p()
    for kw, needle in ((dict(num_samples=0, horizon=5, world_size=1), b"num_samples"),
                       (dict(num_samples=8, horizon=5, world_size=2, rank=2), b"rank"),
                       (dict(num_samples=8, horizon=5, world_size=1), b"sigmas")):
        cfg = _cabi.MppiCfg(**kw)
        assert lib.bnv_mppi_create(C.byref(h), C.byref(cfg)) == -1
        assert needle in lib.bnv_last_error()
    assert lib.bnv_mppi_create(None, None) == -1
    assert lib.bnv_mppi_forward(None, None, None, None, None, None) == -1
    assert lib.bnv_mppi_launch_count(None) == 0


def test_header_is_valid_c_and_a_plain_c_client_links():
    """include/bnv_mppi.h compiles as C99 (no C++-isms behind the extern "C" guard) and a C program links against the
    shared library and reaches it (examples/c_abi_demo.c)."""
    import shutil
    import tempfile

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    build.ensure_built()
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "c_abi_demo")
        cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
               os.path.join(root, "examples", "c_abi_demo.c"), "-L", build.LIB_DIR, "-lbnvmppi",
               "-Wl,-rpath," + build.LIB_DIR, "-o", exe]
        out = subprocess.run(cmd, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        run = subprocess.run([exe], capture_output=True, text=True)
        assert run.returncode == 0, run.stdout + run.stderr
        assert f"bnv_abi_version = {_cabi.ABI_VERSION}" in run.stdout
        if not torch.cuda.is_available():
            assert "bnv_mppi_create -> -2" in run.stdout  # BNV_ERR_CUDA: fails loudly without a device
