"""Host-side mirror of the reference ``DWA`` planner (src/planners/local_planners/dwa.py:17-299) on the sm_100a
rollout kernel.

DWA is the MPPI rollout/cost machinery with K = num_lin_vel * num_ang_vel constant action sequences, no noise, no
control cost, weights = softmax(-cost) and an argmin instead of the weighted mean -- so the engine runs it as an MPPI
handle with lambda_ = 1, a zero mean sequence that is never updated, and the held actions injected as "noise".
Every numeric step is a CUDA kernel of libbnvmppi.so: the dynamic window / linspace / cartesian product
(``bnv_dwa_actions``), the sub-goal selection (``bnv_mppi_dwa_subgoal``), the rollouts + costs + softmax
(``bnv_mppi_forward``), the argmin and gathers (``bnv_mppi_argmin``) and the descending sort of ``get_top_samples``
(``bnv_mppi_top_samples`` with n = K).  Same constructor signature, ``forward`` / ``update_reference_path`` /
``get_top_samples`` contract and attribute names as the reference class.
"""

from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch
import torch.nn as nn

from . import _cabi
from .mppi import _DevView, _introspect_problem


class DWA(nn.Module):
    """Dynamic Window Approach (Fox et al., 1997) on one B200."""

    def __init__(self, horizon: int, dim_state: int, dim_control: int, dynamics, objectives, a_lim: torch.Tensor,
                 delta_t: float, lookahead_distance: float = 1.0, num_lin_vel: int = 10, num_ang_vel: int = 10,
                 device=torch.device("cuda"), dtype=torch.float32, seed: int = 42) -> None:
        super().__init__()
        torch.manual_seed(seed)  # dwa.py:59
        assert dynamics.min_action.shape == (dim_control,), "minimum actions must be a tensor of shape (dim_control,)"
        assert dynamics.max_action.shape == (dim_control,), "maximum actions must be a tensor of shape (dim_control,)"
        assert a_lim.shape == (dim_control,), "acceleration limits must be a tensor of shape (dim_control,)"
        if dim_state != 3 or dim_control != 2:
            raise ValueError("the engine implements the unicycle model: dim_state=3, dim_control=2")
        if dtype != torch.float32:
            raise ValueError("the engine computes in float32 (the reference's default dtype)")
        dev = torch.device(device)
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("benchnav_b200.DWA runs on CUDA (sm_100a) only; there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self._device, self._dtype, self._lib = dev, dtype, _cabi.load()
        self._horizon, self._dim_state, self._dim_control = int(horizon), dim_state, dim_control
        self._dynamics, self._objectives = dynamics, objectives
        self._delta_t, self._lookahead_distance = float(delta_t), float(lookahead_distance)
        self._num_lin_vel, self._num_ang_vel = int(num_lin_vel), int(num_ang_vel)
        self._num_actions = K = self._num_lin_vel * self._num_ang_vel
        lo, hi, al = (t.detach().cpu().to(torch.float32).tolist() for t in (dynamics.min_action, dynamics.max_action, a_lim))
        self._u_min, self._u_max, self._a_lim = (C.c_float * 2)(*lo), (C.c_float * 2)(*hi), (C.c_float * 2)(*al)

        risks, g, res, x_lim, y_lim, goal, thr, dt = _introspect_problem(dynamics, objectives)
        # the rollouts use transit's default time step (dwa.py:205-207 never pass delta_t); delta_t only sizes the window
        cfg = _cabi.MppiCfg(num_samples=K, horizon=self._horizon, lambda_=1.0, dt=dt, seed=int(seed), rank=0, world_size=1,
                            device=dev.index, flags=_cabi.BNV_FLAG_RECORD_STATES)
        for i in range(2):
            cfg.sigma[i], cfg.u_min[i], cfg.u_max[i] = 1.0, lo[i], hi[i]
        self._handle = _cabi.SolverHandle(self._lib, cfg)
        _cabi.check(self._lib.bnv_mppi_set_keep_mean(self._handle, 0))
        self._risk_dev = risks.detach().to(dev, torch.float32).contiguous()
        goal_xy = torch.as_tensor(goal).detach().to("cpu", torch.float32).reshape(-1)[:2].tolist()
        self._goal_host = (C.c_float * 2)(*goal_xy)
        with torch.cuda.device(dev):
            _cabi.check(self._lib.bnv_mppi_set_problem(
                self._handle, self._risk_dev.data_ptr(), g, self._risk_dev.stride(0), res, x_lim[0], x_lim[1],
                y_lim[0], y_lim[1], self._goal_host, thr, self._stream()))
        T = self._horizon
        view = lambda ptr, shape: torch.as_tensor(_DevView(ptr, shape, self._handle), device=dev)  # noqa: E731
        self._weights = view(self._lib.bnv_mppi_weights(self._handle), (K,))
        self._costs = view(self._lib.bnv_mppi_costs(self._handle), (K,))
        self._state_seq_batch = view(self._lib.bnv_mppi_states(self._handle), (K, T + 1, 3))
        self._previous_action_seq = torch.zeros(T, 2, device=dev, dtype=dtype)  # dwa.py:96-98
        self._actions = torch.empty(K, 2, device=dev, dtype=dtype)
        self._controls = torch.empty(K, T, 2, device=dev, dtype=dtype)
        self._sub_goal = torch.empty(2, device=dev, dtype=dtype)
        self._state_dev = torch.zeros(3, device=dev, dtype=dtype)
        self._scratch_u = torch.empty(T, 2, device=dev, dtype=dtype)       # MPPI's weighted mean: unused by DWA
        self._scratch_opt = torch.empty(T + 1, 3, device=dev, dtype=dtype)
        self.reference_path = None

    def _stream(self) -> int:
        return torch.cuda.current_stream(self._device).cuda_stream

    def close(self) -> None:
        """Destroy the engine handle now (device buffers, pinned staging, streams).  Without it the handle lives until
        the solver AND every tensor view of its buffers (``_weights``, ``_state_seq_batch``, ...) are gone."""
        self._handle.close()

    def update_reference_path(self, reference_path: torch.Tensor) -> None:
        """dwa.py:151-158."""
        if reference_path is not None:
            assert reference_path.shape[1] == 2, "reference_path must be a tensor of shape (num_positions, 2)"
            self.reference_path = reference_path.detach().to(self._device, self._dtype).contiguous()

    def forward(self, state: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """dwa.py:116-149: returns ``(optimal_action_seq [1,2], optimal_state_seq [1,T+1,3])``."""
        if not torch.is_tensor(state):
            state = torch.tensor(state, dtype=self._dtype)
        assert state.shape == (self._dim_state,), "state must be a tensor of shape (dim_state,)"
        lib, h, s = self._lib, self._handle, self._stream()
        with torch.cuda.device(self._device):
            self._state_dev.copy_(state.detach().to(self._dtype), non_blocking=True)
            prev = self._previous_action_seq  # [T,2] zeros at first, then the [1,2] optimum (dwa.py:146): row 0 either way
            _cabi.check(lib.bnv_dwa_actions(prev.data_ptr(), self._u_min, self._u_max, self._a_lim, self._delta_t,
                                            self._num_lin_vel, self._num_ang_vel, self._horizon,
                                            self._actions.data_ptr(), self._controls.data_ptr(), s))
            if self.reference_path is not None:
                _cabi.check(lib.bnv_mppi_dwa_subgoal(h, self.reference_path.data_ptr(), int(self.reference_path.shape[0]),
                                                     self._state_dev.data_ptr(), self._actions.data_ptr(),
                                                     self._lookahead_distance, self._sub_goal.data_ptr(), s))
                _cabi.check(lib.bnv_mppi_set_goal_dev(h, self._sub_goal.data_ptr()))
            else:
                _cabi.check(lib.bnv_mppi_set_goal_dev(h, None))
            _cabi.check(lib.bnv_mppi_forward(h, self._state_dev.data_ptr(), self._controls.data_ptr(),
                                             self._scratch_u.data_ptr(), self._scratch_opt.data_ptr(), s))
            action = torch.empty(1, 2, device=self._device, dtype=self._dtype)
            states = torch.empty(1, self._horizon + 1, 3, device=self._device, dtype=self._dtype)
            _cabi.check(lib.bnv_mppi_argmin(h, self._actions.data_ptr(), action.data_ptr(), states.data_ptr(), None, s))
        self._previous_action_seq = action  # dwa.py:146
        return action, states

    def get_top_samples(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """dwa.py:287-299: every rollout, sorted by descending weight."""
        K = self._num_actions
        states = torch.empty(K, self._horizon + 1, 3, device=self._device, dtype=self._dtype)
        weights = torch.empty(K, device=self._device, dtype=self._dtype)
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_mppi_top_samples(self._handle, K, states.data_ptr(), weights.data_ptr(), self._stream()))
        return states, weights

    @property
    def launch_count(self) -> int:
        return int(self._lib.bnv_mppi_launch_count(self._handle))
