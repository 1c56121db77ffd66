"""``BatchedMPPI``: E independent MPPI planners -- one per ``PlanetaryEnv`` instance -- solved in ONE kernel launch
(BASELINE config 3: "batched 64 planetary_env instances x K=4096, T=30").

The reference has no batched planner: config 3's meaning is "E separate reference ``MPPI`` objects stepped in a Python
loop" (SURVEY 8c), which is also how the parity tests check this class.  Here every per-solver buffer of the engine
simply carries a leading E and the rollout kernel's grid gains a y dimension (environment); each environment has its
own risk map, state, goal and mean sequence.  Environments, not samples, are what shards across GPUs: give every rank
its own slice of the environment list (``benchnav_b200.dist.shard_range(E, rank, world)``) -- there is nothing to
exchange, so no collective is involved.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _cabi
from .mppi import _DevView, _introspect_problem


class BatchedMPPI(nn.Module):
    """E reference-shaped (dynamics, objectives) pairs with a common grid geometry -> one batched solver."""

    def __init__(self, horizon: int, num_samples: int, dynamics_list: Sequence, objectives_list: Sequence,
                 sigmas: torch.Tensor, lambda_: float, device=torch.device("cuda"), seed: int = 42) -> None:
        super().__init__()
        if len(dynamics_list) != len(objectives_list) or len(dynamics_list) < 1:
            raise ValueError("need one objectives object per dynamics object")
        assert sigmas.shape == (2,), "sigmas must be a tensor of shape (dim_control,)"
        dev = torch.device(device)
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("benchnav_b200.BatchedMPPI runs on CUDA (sm_100a) only; there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self._device = dev
        self._num_envs = E = len(dynamics_list)
        self._horizon, self._num_samples = int(horizon), int(num_samples)
        self._dynamics, self._objectives = list(dynamics_list), list(objectives_list)
        self._lib = _cabi.load()
        infos = [_introspect_problem(d, o) for d, o in zip(dynamics_list, objectives_list)]
        _, g, res, x_lim, y_lim, _, thr, dt = infos[0]
        for info in infos[1:]:
            if (info[1], info[2], info[3], info[4], info[6], info[7]) != (g, res, x_lim, y_lim, thr, dt):
                raise ValueError("all environments of a batch must share grid geometry, threshold and time step")
        d0 = dynamics_list[0]
        for d in dynamics_list[1:]:
            if not (torch.equal(d.min_action.cpu(), d0.min_action.cpu()) and torch.equal(d.max_action.cpu(), d0.max_action.cpu())):
                raise ValueError("all environments of a batch must share the action bounds")
        cfg = _cabi.MppiCfg(num_samples=self._num_samples, horizon=self._horizon, lambda_=float(lambda_), dt=dt,
                            seed=int(seed) & 0xFFFFFFFFFFFFFFFF, rank=0, world_size=1, device=dev.index,
                            flags=_cabi.BNV_FLAG_RECORD_STATES, num_envs=E)
        sig, lo, hi = (t.detach().cpu().to(torch.float32).tolist() for t in (sigmas, d0.min_action, d0.max_action))
        for i in range(2):
            cfg.sigma[i], cfg.u_min[i], cfg.u_max[i] = sig[i], lo[i], hi[i]
        self._handle = _cabi.SolverHandle(self._lib, cfg)
        self._geom = (g, res, x_lim, y_lim, thr)
        self.sync_problems()
        k, t = self._num_samples, self._horizon
        view = lambda ptr, shape: torch.as_tensor(_DevView(ptr, shape, self._handle), device=dev)  # noqa: E731
        self._weights = view(self._lib.bnv_mppi_weights(self._handle), (E, k))
        self._costs = view(self._lib.bnv_mppi_costs(self._handle), (E, k))
        self._previous_action_seq = view(self._lib.bnv_mppi_u_prev(self._handle), (E, t, 2))
        self._action_noises = view(self._lib.bnv_mppi_noise(self._handle), (E, k, t, 2))
        self._state_seq_batch = view(self._lib.bnv_mppi_states(self._handle), (E, k, t + 1, 3))

    def _stream(self) -> int:
        return torch.cuda.current_stream(self._device).cuda_stream

    def sync_problems(self) -> None:
        """(Re)upload every environment's risk map and goal (call again after changing them)."""
        g, res, x_lim, y_lim, thr = self._geom
        risks = torch.stack([d._traversability_model._risks.detach().to(torch.float32).cpu() for d in self._dynamics])
        self._risk_dev = risks.to(self._device).contiguous()
        goals = (C.c_float * (2 * self._num_envs))()
        for e, o in enumerate(self._objectives):
            gx, gy = torch.as_tensor(o._goal_pos).detach().to("cpu", torch.float32).reshape(-1)[:2].tolist()
            goals[2 * e], goals[2 * e + 1] = gx, gy
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_mppi_set_problem_ex(
                self._handle, self._risk_dev.data_ptr(), None, g, self._risk_dev.stride(1), self._risk_dev.stride(0),
                res, x_lim[0], x_lim[1], y_lim[0], y_lim[1], goals, thr, self._stream()))

    def close(self) -> None:
        """Destroy the engine handle now (device buffers, pinned staging, streams).  Without it the handle lives until
        the solver AND every tensor view of its buffers (``_weights``, ``_state_seq_batch``, ...) are gone."""
        self._handle.close()

    def forward(self, states: torch.Tensor, noise: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """One control iteration of every environment (mppi.py:130-219 x E).

        ``states`` [E,3] -> ``(optimal_action_seq [E,T,2], optimal_state_seq [E,1,T+1,3])``;
        ``noise`` (optional, [E,K,T,2]) injects the sigma-scaled control noise."""
        E, t, k = self._num_envs, self._horizon, self._num_samples
        assert tuple(states.shape) == (E, 3)
        states = states.detach().to(self._device, torch.float32).contiguous()
        noise_ptr = None
        if noise is not None:
            if tuple(noise.shape) != (E, k, t, 2):
                raise ValueError(f"noise must have shape {(E, k, t, 2)}")
            noise = noise.detach().to(self._device, torch.float32).contiguous()
            noise_ptr = noise.data_ptr()
        u_opt = torch.empty(E, t, 2, device=self._device, dtype=torch.float32)
        opt_states = torch.empty(E, 1, t + 1, 3, device=self._device, dtype=torch.float32)
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_mppi_forward(self._handle, states.data_ptr(), noise_ptr, u_opt.data_ptr(),
                                                   opt_states.data_ptr(), self._stream()))
        self._keepalive = (states, noise)
        return u_opt, opt_states

    def forward_host(self, states: torch.Tensor, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
        """``forward`` for HOST states [E,3], returning HOST tensors ``(u* [E,T,2], optimal states [E,1,T+1,3])``: one
        staged copy each way and one synchronisation inside the library (``bnv_mppi_forward_host_batch``) instead of
        tensor copies around ``forward``.  ``out``: optional caller-owned contiguous fp32 CPU tensors."""
        E, t = self._num_envs, self._horizon
        if not (torch.is_tensor(states) and states.dtype == torch.float32 and states.device.type == "cpu"
                and states.is_contiguous()):
            states = torch.as_tensor(states, dtype=torch.float32).detach().cpu().contiguous()
        assert tuple(states.shape) == (E, 3)
        if out is None:
            u_opt = torch.empty(E, t, 2, dtype=torch.float32)
            opt_states = torch.empty(E, 1, t + 1, 3, dtype=torch.float32)
        else:
            u_opt, opt_states = out
            if not (u_opt.dtype == torch.float32 and opt_states.dtype == torch.float32 and u_opt.device.type == "cpu"
                    and opt_states.device.type == "cpu" and u_opt.is_contiguous() and opt_states.is_contiguous()
                    and tuple(u_opt.shape) == (E, t, 2) and tuple(opt_states.shape) == (E, 1, t + 1, 3)):
                raise ValueError("out must be contiguous fp32 CPU tensors of shapes [E,T,2] and [E,1,T+1,3]")
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_mppi_forward_host_batch(self._handle, states.data_ptr(), u_opt.data_ptr(),
                                                              opt_states.data_ptr(), self._stream()))
        return u_opt, opt_states

    def get_top_samples(self, num_samples: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Top ``num_samples`` rollouts of every environment by weight (mppi.py:221-240 x E): [E,n,T+1,3], [E,n]."""
        assert num_samples <= self._num_samples
        n, E = int(num_samples), self._num_envs
        states = torch.empty(E, n, self._horizon + 1, 3, device=self._device, dtype=torch.float32)
        weights = torch.empty(E, n, device=self._device, dtype=torch.float32)
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_mppi_top_samples(self._handle, n, states.data_ptr(), weights.data_ptr(),
                                                       self._stream()))
        return states, weights

    def reset(self) -> None:
        _cabi.check(self._lib.bnv_mppi_reset(self._handle, self._stream()))

    def graph_capturable(self, enable: bool = True, external_advance: bool = False) -> None:
        """Keep the iteration counter (Philox counter word, launch epoch) in device memory so that ``forward`` can be
        captured in a CUDA graph (``torch.cuda.graph``) and replayed: one graph launch per control step.
        ``external_advance``: the counter is advanced by the caller's own kernel (``BatchedPlanetaryEnv.closed_loop_step``
        does it) instead of a one-thread kernel behind every ``forward``."""
        with torch.cuda.device(self._device):
            mode = (2 if external_advance else 1) if enable else 0
            _cabi.check(self._lib.bnv_mppi_device_counter(self._handle, mode, self._stream()))

    @property
    def iteration_counter_ptr(self) -> int:
        """Device address of the iteration counter in graph-capturable mode (0 otherwise)."""
        return int(self._lib.bnv_mppi_iteration_counter(self._handle) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.bnv_mppi_launch_count(self._handle))

    @property
    def launch_geometry(self) -> dict:
        """Rollout-kernel launch shape: CTAs per environment, warps per CTA, slab split step, cooperative."""
        out = (C.c_int32 * 4)()
        _cabi.check(self._lib.bnv_mppi_launch_geometry(self._handle, out))
        return {"ctas": out[0], "envs": self._num_envs, "warps_per_cta": out[1], "rec_split": out[2],
                "cooperative": bool(out[3])}

    def kernel_timing(self, max_launches: int) -> None:
        """Record CUDA-event pairs around the rollout kernel of the next ``max_launches`` iterations."""
        _cabi.check(self._lib.bnv_mppi_kernel_timing(self._handle, int(max_launches)))

    def kernel_time(self) -> Tuple[float, int]:
        """(summed rollout-kernel milliseconds, launches measured) since kernel_timing(); synchronises."""
        ms, n = C.c_double(), C.c_uint64()
        _cabi.check(self._lib.bnv_mppi_kernel_time(self._handle, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)
