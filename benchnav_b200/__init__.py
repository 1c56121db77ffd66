"""benchnav_b200 -- a Blackwell-native (sm_100a) MPPI engine behind BenchNav's planner API.

Public surface:
    MPPI                 drop-in for the reference ``src.planners.local_planners.mppi.MPPI``
    BatchedMPPI          E planners (one per environment) in one launch (BASELINE config 3)
    DWA                  drop-in for ``src.planners.local_planners.dwa.DWA`` on the same rollout kernel
    BatchedPlanetaryEnv  ``PlanetaryEnv.step / collision_check`` for E environments on the device
    infer_risk_map       ``TraversabilityModel._infer_risk_map`` on the device
    build_library()      (re)compile libbnvmppi.so in-tree with nvcc
    synthetic            synthetic terrain / problem generators used by bench.py and the tests
"""

from .build import build_library, ensure_built  # noqa: F401


def __getattr__(name):
    # import torch-dependent pieces lazily so that `import benchnav_b200` stays cheap
    if name == "MPPI":
        from .mppi import MPPI

        return MPPI
    if name == "BatchedMPPI":
        from .batch import BatchedMPPI

        return BatchedMPPI
    if name == "DWA":
        from .dwa import DWA

        return DWA
    if name == "BatchedPlanetaryEnv":
        from .env import BatchedPlanetaryEnv

        return BatchedPlanetaryEnv
    if name == "infer_risk_map":
        from .risk import infer_risk_map

        return infer_risk_map
    raise AttributeError(name)
