"""benchnav_b200 -- a Blackwell-native (sm_100a) MPPI engine behind BenchNav's planner API.

Public surface:
    MPPI                 drop-in for the reference ``src.planners.local_planners.mppi.MPPI``
    build_library()      (re)compile libbnvmppi.so in-tree with nvcc
    synthetic            synthetic terrain / problem generators used by bench.py and the tests
"""

from .build import build_library, ensure_built  # noqa: F401


def __getattr__(name):
    # import torch-dependent pieces lazily so that `import benchnav_b200` stays cheap
    if name == "MPPI":
        from .mppi import MPPI

        return MPPI
    raise AttributeError(name)
