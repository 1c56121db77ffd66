"""In-tree build of libbnvmppi.so (the C-ABI CUDA library) for sm_100a with nvcc.

The library has no PyTorch dependency: it is compiled straight from ``csrc/*.cu`` and loaded with
ctypes (``_cabi.py``).  The built ``.so`` is git-ignored but travels to the GPU box with the repo
snapshot.  ``python -m benchnav_b200.build`` rebuilds it; ``ensure_built()`` rebuilds only when a
source is newer than the library.
"""

from __future__ import annotations

import fcntl
import os
import shutil
import subprocess
import sys
import warnings

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libbnvmppi.so")

COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]


SOURCES = ("bnv_mppi.cu", "rollout_ext.cu", "rollout_wide.cu", "rollout_wide_ext.cu", "bnv_aux.cu")


def _sources():
    srcs = [os.path.join(CSRC, f) for f in SOURCES]
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(INCLUDE, "bnv_mppi.h")]
    return srcs, deps


def nvcc_path() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libbnvmppi.so cannot be built")
    return exe


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    _, deps = _sources()
    lib_m = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > lib_m for d in deps)


def build_library(verbose: bool = False) -> str:
    """Compile every translation unit (in parallel: the rollout kernel's template space dominates) and link.

    Safe against concurrent callers (every rank of a torchrun job may find the library stale at once): the build runs
    under an exclusive file lock, objects go to a per-process directory and the library is linked to a temporary
    path and moved into place atomically, so no process ever loads a half-written file."""
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    srcs, _ = _sources()
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = nvcc_path()
    procs = []
    for src in srcs:
        obj = os.path.join(obj_dir, os.path.splitext(os.path.basename(src))[0] + ".o")
        cmd = [nvcc, *COMPILE_FLAGS, "-I", INCLUDE, "-I", CSRC, "-c", "-o", obj, src]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log, objs = "", []
    for cmd, obj, proc in procs:
        out, _ = proc.communicate()
        log += " ".join(cmd) + "\n" + out
        if proc.returncode != 0:
            for _, _, other in procs:
                if other.poll() is None:
                    other.kill()
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
        objs.append(obj)
    tmp_lib = LIB_PATH + f".tmp{os.getpid()}"
    link = [nvcc, *LINK_FLAGS, "-o", tmp_lib, *objs]
    proc = subprocess.run(link, capture_output=True, text=True)
    log += " ".join(link) + "\n" + proc.stdout + proc.stderr
    if proc.returncode != 0:
        if os.path.exists(tmp_lib):
            os.remove(tmp_lib)
        raise RuntimeError("link failed:\n" + " ".join(link) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp_lib, LIB_PATH)
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write(log)
    if verbose:
        print(log)
    return LIB_PATH


def ensure_built() -> str:
    """Build if missing or stale.  A stale library that cannot be rebuilt is an error when nvcc is present (the sources
    changed and do not compile); without nvcc (a deployment box that received a prebuilt library) it is loaded with a
    loud warning -- the ABI-version check in ``_cabi.load`` is then the only guard."""
    if is_stale():
        have_nvcc = bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")
        if have_nvcc:
            os.makedirs(LIB_DIR, exist_ok=True)
            with open(os.path.join(LIB_DIR, ".build.lock"), "a") as lock:  # another rank may be building right now
                fcntl.flock(lock, fcntl.LOCK_EX)
                fcntl.flock(lock, fcntl.LOCK_UN)
            if is_stale():
                build_library()
        elif not os.path.exists(LIB_PATH):
            raise RuntimeError("libbnvmppi.so is missing and nvcc is not available to build it")
        else:
            warnings.warn("libbnvmppi.so is older than its sources and nvcc is not available to rebuild it: "
                          "loading the existing library", RuntimeWarning, stacklevel=2)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(verbose="-v" in sys.argv))
