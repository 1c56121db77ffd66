"""Host-side mirror of the reference planner class, running on the sm_100a engine.

``MPPI`` keeps the constructor signature and the ``forward`` / ``get_top_samples`` contract of the
reference ``MPPI(nn.Module)`` (src/planners/local_planners/mppi.py:17-240) so that Tutorial 3.3 and
``test/test_mppi.py:160-185`` run unchanged, and reads -- never replaces -- the reference's own
``UnicycleModel`` / ``Objectives`` / ``GridMap`` objects (duck-typed: anything exposing the same
attributes works).  All numerics happen in ``libbnvmppi.so`` through the C ABI (``_cabi.py``); PyTorch is
used for device memory, streams and ``torch.distributed`` only.  There is no CPU path: constructing the
solver without a CUDA device raises.

Differences from the reference that a caller can observe (all documented in DESIGN.md):
  * ``device`` must name a CUDA device; the reference silently falls back to CPU (mppi.py:69-72).
  * the noise stream: ``noise_source="philox"`` (default) draws inside the engine from a counter-based
    generator keyed by (seed, global sample, iteration); ``noise_source="torch"`` draws with
    ``torch.empty(K,T,2).normal_()`` on the CUDA generator exactly like the reference's
    ``MultivariateNormal.rsample`` on CUDA (mppi.py:149-151), throw-away constructor draw included;
    ``forward(state, noise=...)`` injects the sigma-scaled noise (parity tests).
  * ``solve`` is an alias of ``forward`` (north_star names it; the reference has only ``forward``).
"""

from __future__ import annotations

import ctypes as C
import inspect
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _cabi
from .dist import ShardInfo, attach_peer_mailboxes, gather_shard_partials, merge_top_candidates


class _DevView:
    """Zero-copy view of an engine-owned device buffer via ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], owner) -> None:
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner  # keep the handle alive while the view exists


def _slip_distribution(grid_map, key: str = "predictions"):
    """(mean, std) tensors of ``GridMap.distributions[key]`` (grid_map.py:24-33): what the observation-mode lookup
    samples from (traversability_model.py:65-69)."""
    try:
        dist = grid_map.distributions[key]
        mean, std = dist.mean, dist.stddev
    except (AttributeError, KeyError, TypeError) as exc:
        raise TypeError(f"grid map has no Normal distribution under distributions['{key}']") from exc
    g = int(grid_map.grid_size)
    if tuple(mean.shape) != (g, g) or tuple(std.shape) != (g, g):
        raise ValueError("slip distribution mean/stddev must be [grid_size, grid_size]")
    return mean, std


def _introspect_problem(dynamics, objectives):
    """Pull what ``forward`` reads through ``dynamics``/``objectives`` out of the objects."""
    try:
        grid_map = dynamics._grid_map
        model_config = dynamics._model_config
        risks = dynamics._traversability_model._risks
        goal = objectives._goal_pos
        thr = objectives._stuck_threshold
        g, res = int(grid_map.grid_size), float(grid_map.resolution)
        x_lim, y_lim = tuple(grid_map.x_limits), tuple(grid_map.y_limits)
    except AttributeError as exc:  # not UnicycleModel + Objectives shaped
        raise TypeError("benchnav_b200.MPPI needs UnicycleModel-like dynamics and Objectives-like objectives "
                        f"(missing attribute: {exc})") from exc
    if getattr(model_config, "mode", None) != "inference":
        # observation mode makes transit return a tuple and breaks the reference as well (robot_model.py:97-98)
        raise ValueError("dynamics must use ModelConfig(mode='inference')")
    if risks is None or not torch.is_tensor(risks) or risks.dim() != 2 or risks.shape[0] != g or risks.shape[1] != g:
        raise ValueError("dynamics._traversability_model._risks must be a [grid_size, grid_size] tensor")
    dt = 0.1
    try:
        dflt = inspect.signature(dynamics.transit).parameters["delta_t"].default
        if isinstance(dflt, (int, float)):
            dt = float(dflt)  # MPPI never passes delta_t (mppi.py:163): rollouts use transit's default
    except (KeyError, TypeError, ValueError, AttributeError):
        pass
    return risks, g, res, x_lim, y_lim, goal, float(thr), dt


class MPPI(nn.Module):
    """Model Predictive Path Integral control (Williams et al., T-RO 2017) on one or more B200s."""

    def __init__(self, horizon: int, num_samples: int, dim_state: int, dim_control: int, dynamics, objectives,
                 sigmas: torch.Tensor, lambda_: float, device=torch.device("cuda"), dtype=torch.float32,
                 seed: int = 42, *, noise_source: str = "philox", record_states: bool = True,
                 process_group=None, exchange: str = "p2p", stochastic_slip: bool = False) -> None:
        super().__init__()
        torch.manual_seed(seed)  # mppi.py:55
        # same shape checks (and exception type) as mppi.py:58-66
        assert dynamics.min_action.shape == (dim_control,), "minimum actions must be a tensor of shape (dim_control,)"
        assert dynamics.max_action.shape == (dim_control,), "maximum actions must be a tensor of shape (dim_control,)"
        assert sigmas.shape == (dim_control,), "sigmas must be a tensor of shape (dim_control,)"
        if dim_state != 3 or dim_control != 2:
            raise ValueError("the engine implements the unicycle model: dim_state=3, dim_control=2")
        if dtype != torch.float32:
            raise ValueError("the engine computes in float32 (the reference's default dtype)")
        if noise_source not in ("philox", "torch"):
            raise ValueError("noise_source must be 'philox' or 'torch'")
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' (fused, NVLink peer memory) or 'nccl' (all-gather + finalize)")
        dev = torch.device(device)
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("benchnav_b200.MPPI runs on CUDA (sm_100a) only; there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self._device, self._dtype = dev, dtype
        self._horizon, self._num_samples = int(horizon), int(num_samples)
        self._dim_state, self._dim_control = dim_state, dim_control
        self._dynamics, self._objectives = dynamics, objectives
        self._lambda = float(lambda_)
        self._noise_source = noise_source
        self._u_min = dynamics.min_action.clone().detach().to(dev, dtype)
        self._u_max = dynamics.max_action.clone().detach().to(dev, dtype)
        self._sigmas = sigmas.clone().detach().to(dev, dtype)
        self._shard = ShardInfo.from_group(process_group)
        self._lib = _cabi.load()
        # BASELINE config 4: every lookup of the rollouts samples the slip prediction of the cell
        self._stochastic = bool(stochastic_slip)
        if self._stochastic and self._shard.world_size != 1:
            raise ValueError("stochastic_slip needs a single-GPU solver")
        if self._stochastic and noise_source != "philox":
            raise ValueError("stochastic_slip draws its noise in the engine (noise_source='philox')")

        risks, g, res, x_lim, y_lim, goal, thr, dt = _introspect_problem(dynamics, objectives)
        cfg = _cabi.MppiCfg(num_samples=self._num_samples, horizon=self._horizon, lambda_=self._lambda, dt=dt,
                            seed=int(seed) & 0xFFFFFFFFFFFFFFFF, rank=self._shard.rank,
                            world_size=self._shard.world_size, device=dev.index,
                            flags=(_cabi.BNV_FLAG_RECORD_STATES if record_states else 0)
                            | (_cabi.BNV_FLAG_STOCHASTIC_SLIP if self._stochastic else 0))
        sig, lo, hi = (t.detach().cpu().to(torch.float32).tolist() for t in (sigmas, dynamics.min_action, dynamics.max_action))
        for i in range(2):
            cfg.sigma[i], cfg.u_min[i], cfg.u_max[i] = sig[i], lo[i], hi[i]
        self._handle = _cabi.SolverHandle(self._lib, cfg)
        record_states = bool(record_states) or self._stochastic
        self._record_states = record_states
        self._local_samples = int(self._lib.bnv_mppi_local_samples(self._handle))
        self._sample_offset = int(self._lib.bnv_mppi_sample_offset(self._handle))
        self._sample_shape = torch.Size([self._local_samples, self._horizon])
        self._risk_dev: Optional[torch.Tensor] = None
        self._risk_key = None
        self._goal_host = (C.c_float * 2)()
        self._sync_problem(force=True)

        # engine-owned state exposed under the reference's attribute names (zero-copy views)
        k, t = self._local_samples, self._horizon
        view = lambda ptr, shape: torch.as_tensor(_DevView(ptr, shape, self._handle), device=dev)  # noqa: E731
        self._weights = view(self._lib.bnv_mppi_weights(self._handle), (k,))
        self._costs = view(self._lib.bnv_mppi_costs(self._handle), (k,))
        self._previous_action_seq = view(self._lib.bnv_mppi_u_prev(self._handle), (t, 2))
        self._engine_noise = view(self._lib.bnv_mppi_noise(self._handle), (k, t, 2))
        self._state_seq_batch = (view(self._lib.bnv_mppi_states(self._handle), (k, t + 1, 3))
                                 if record_states else None)
        self._action_noises = self._engine_noise
        if noise_source == "torch":
            self._action_noises = self._draw_torch_noise()  # the reference's throw-away draw, mppi.py:105-107
        plen = int(self._lib.bnv_mppi_partial_len(self._handle))
        self._partial = view(self._lib.bnv_mppi_partial(self._handle), (plen,))
        self._gathered = (torch.empty(self._shard.world_size, plen, device=dev, dtype=torch.float32)
                          if self._shard.world_size > 1 else None)
        self._state_dev = torch.zeros(3, device=dev, dtype=torch.float32)
        self._forward_host_fn = self._lib.bnv_mppi_forward_host
        self._forward_action_fn = self._lib.bnv_mppi_forward_host_action
        self._pending_states = None
        self._fused_exchange = False
        if self._shard.world_size > 1 and exchange == "p2p":
            self._fused_exchange = attach_peer_mailboxes(self._lib, self._handle, self._shard)
        # the host-buffer call goes straight to the engine for an unsharded solver, and for a sharded one whose ranks
        # exchange inside the kernel (then this rank leads and the others call forward_follow)
        self._slow_host_path = noise_source != "philox" or (self._shard.world_size != 1 and not self._fused_exchange)

    # ------------------------------------------------------------------ plumbing
    def _draw_torch_noise(self) -> torch.Tensor:
        """MultivariateNormal(0, diag(sigma^2)).rsample on CUDA == sigma * empty(K,T,2).normal_() (mppi.py:149-151)."""
        k, t = self._num_samples, self._horizon
        eps = torch.empty(k, t, 2, device=self._device, dtype=torch.float32).normal_()
        eps = eps[self._sample_offset:self._sample_offset + self._local_samples]
        return (eps * self._sigmas).contiguous()

    def _stream(self) -> int:
        try:  # raw handle of torch's current stream without building a Stream object (~1 us saved per call)
            return torch._C._cuda_getCurrentRawStream(self._device.index)
        except AttributeError:
            return torch.cuda.current_stream(self._device).cuda_stream

    def _sync_problem(self, force: bool = False) -> None:
        """(Re)upload the risk map / goal / threshold when the reference objects changed (cheap identity check)."""
        dyn, obj = self._dynamics, self._objectives
        risks, goal = dyn._traversability_model._risks, obj._goal_pos
        if self._stochastic:
            risks = _slip_distribution(dyn._grid_map)[0]
        quick = (id(risks), risks._version, id(goal), goal._version if torch.is_tensor(goal) else None,
                 obj._stuck_threshold, id(dyn._grid_map))
        if not force and quick == self._risk_key:
            return
        risks, g, res, x_lim, y_lim, goal, thr, _ = _introspect_problem(dyn, obj)
        goal_xy = torch.as_tensor(goal).detach().to("cpu", torch.float32).reshape(-1)[:2].tolist()
        self._goal_host[0], self._goal_host[1] = goal_xy
        with torch.cuda.device(self._device):
            if self._stochastic:
                mean, std = _slip_distribution(dyn._grid_map)
                self._risk_dev = mean.detach().to(self._device, torch.float32).contiguous()
                self._std_dev = std.detach().to(self._device, torch.float32).contiguous()
                _cabi.check(self._lib.bnv_mppi_set_problem_ex(
                    self._handle, self._risk_dev.data_ptr(), self._std_dev.data_ptr(), g, self._risk_dev.stride(0), 0,
                    res, x_lim[0], x_lim[1], y_lim[0], y_lim[1], self._goal_host, thr, self._stream()))
            else:
                self._risk_dev = risks.detach().to(self._device, torch.float32).contiguous()
                _cabi.check(self._lib.bnv_mppi_set_problem(
                    self._handle, self._risk_dev.data_ptr(), g, self._risk_dev.stride(0), res, x_lim[0], x_lim[1],
                    y_lim[0], y_lim[1], self._goal_host, thr, self._stream()))
        self._risk_key = quick

    def close(self) -> None:
        """Destroy the engine handle now (device buffers, pinned staging, streams).  Without it the handle lives until
        the solver AND every tensor view of its buffers (``_weights``, ``_state_seq_batch``, ...) are gone."""
        self._handle.close()

    # ------------------------------------------------------------------ reference API
    def forward(self, state: torch.Tensor, noise: Optional[torch.Tensor] = None, xi: Optional[torch.Tensor] = None,
                xi_opt: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """One control iteration (mppi.py:130-219).

        Returns ``(optimal_action_seq [T,2], optimal_state_seq [1,T+1,3])`` on the solver's device.
        ``noise`` (optional, [K_local,T,2]) injects the sigma-scaled control noise of this shard.
        ``xi`` [K,2T+1] / ``xi_opt`` [T] (stochastic_slip only, together with ``noise``) inject the standard normals
        behind the lookups' ``Normal.sample()`` calls (layout: include/bnv_mppi.h, bnv_mppi_forward_ex).
        """
        if not torch.is_tensor(state):
            state = torch.tensor(state, dtype=self._dtype)
        assert state.shape == (self._dim_state,)
        self._sync_problem()
        with torch.cuda.device(self._device):
            state_ptr, state_host = None, None
            if state.device.type == "cuda":
                if state.device == self._device and state.dtype == torch.float32 and state.is_contiguous():
                    state_ptr = state.detach().data_ptr()  # already resident: read in place (stream-ordered)
                else:
                    self._state_dev.copy_(state.detach(), non_blocking=True)
                    state_ptr = self._state_dev.data_ptr()
            else:  # host state: three floats by value in the launch packet, no device copy
                state_host = (C.c_float * 3)(*state.detach().to(torch.float32).tolist())
            if noise is not None:
                if noise.shape != (self._local_samples, self._horizon, 2):
                    raise ValueError(f"noise must have shape {(self._local_samples, self._horizon, 2)}")
                noise = noise.detach().to(self._device, torch.float32).contiguous()
                self._action_noises = noise
            elif self._noise_source == "torch":
                noise = self._action_noises = self._draw_torch_noise()
            else:
                self._action_noises = self._engine_noise
            u_opt = torch.empty(self._horizon, 2, device=self._device, dtype=torch.float32)
            opt_states = torch.empty(1, self._horizon + 1, 3, device=self._device, dtype=torch.float32)
            stream = self._stream()
            noise_ptr = noise.data_ptr() if noise is not None else None
            if self._stochastic:
                if (noise is None) != (xi is None) or (noise is None) != (xi_opt is None):
                    raise ValueError("stochastic_slip: give noise, xi and xi_opt together or none of them")
                if state_ptr is None:
                    self._state_dev.copy_(state.detach().to(torch.float32), non_blocking=True)
                    state_ptr = self._state_dev.data_ptr()
                xi_ptr = xo_ptr = None
                if xi is not None:
                    if tuple(xi.shape) != (self._local_samples, 2 * self._horizon + 1) or tuple(xi_opt.shape) != (self._horizon,):
                        raise ValueError("xi must be [K, 2T+1] and xi_opt [T]")
                    xi = xi.detach().to(self._device, torch.float32).contiguous()
                    xi_opt = xi_opt.detach().to(self._device, torch.float32).contiguous()
                    self._xi_keepalive = (xi, xi_opt)
                    xi_ptr, xo_ptr = xi.data_ptr(), xi_opt.data_ptr()
                _cabi.check(self._lib.bnv_mppi_forward_ex(self._handle, state_ptr, noise_ptr, xi_ptr, xo_ptr,
                                                          u_opt.data_ptr(), opt_states.data_ptr(), stream))
            elif state_host is None:
                _cabi.check(self._lib.bnv_mppi_forward(self._handle, state_ptr, noise_ptr, u_opt.data_ptr(),
                                                       opt_states.data_ptr(), stream))
            else:
                _cabi.check(self._lib.bnv_mppi_forward_state(self._handle, state_host, noise_ptr, u_opt.data_ptr(),
                                                             opt_states.data_ptr(), stream))
            if self._shard.world_size > 1 and not self._fused_exchange:
                gather_shard_partials(self._partial, self._gathered, self._shard)
                _cabi.check(self._lib.bnv_mppi_finalize(self._handle, self._gathered.data_ptr(), u_opt.data_ptr(),
                                                        opt_states.data_ptr(), stream))
        return u_opt, opt_states

    solve = forward  # north_star's name for the same call

    def forward_host(self, state: torch.Tensor, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
        """``forward`` for a HOST state, returning HOST tensors: the ABI's host-buffer form
        (``bnv_mppi_forward_host``): H2D of the state, the iteration, D2H of both results, one completion wait.

        ``out = (u_opt [T,2], opt_states [1,T+1,3])``: optional caller-owned fp32 CPU tensors to write the results into
        (as with the C ABI, where the caller owns every buffer); fresh tensors are allocated otherwise.

        The call is the end-to-end path of a control loop, so its per-call Python work is kept minimal: argument
        objects that were validated on the previous call (same tensor objects) are not validated again."""
        d = self.__dict__  # plain attributes only: bypasses nn.Module's __getattr__/__setattr__ machinery
        if out is not None:
            if out is not d.get("_host_out_ok"):
                u_opt, opt_states = out
                if not (u_opt.dtype == torch.float32 and opt_states.dtype == torch.float32
                        and u_opt.device.type == "cpu" and opt_states.device.type == "cpu" and u_opt.is_contiguous()
                        and opt_states.is_contiguous() and u_opt.shape == (self._horizon, 2)
                        and opt_states.shape == (1, self._horizon + 1, 3)):
                    raise ValueError("out must be contiguous fp32 CPU tensors of shapes [T,2] and [1,T+1,3]")
                d["_host_out_ok"] = out
            else:
                u_opt, opt_states = out
        if d["_slow_host_path"]:
            res = self.forward(state)
            if out is not None:
                u_opt.copy_(res[0])
                opt_states.copy_(res[1])
                return u_opt, opt_states
            return res[0].cpu(), res[1].cpu()
        if state is not d.get("_host_state_ok"):
            if not (torch.is_tensor(state) and state.dtype == torch.float32 and state.device.type == "cpu"
                    and state.is_contiguous()):
                state = torch.as_tensor(state, dtype=torch.float32).detach().cpu().contiguous()
            assert state.shape == (self._dim_state,)
            d["_host_state_ok"] = state
        self._sync_problem()
        if out is None:
            u_opt = torch.empty(self._horizon, 2, dtype=torch.float32)
            opt_states = torch.empty(1, self._horizon + 1, 3, dtype=torch.float32)
        if d["_action_noises"] is not d["_engine_noise"]:
            d["_action_noises"] = d["_engine_noise"]
        rc = d["_forward_host_fn"](d["_handle"], state.data_ptr(), None, u_opt.data_ptr(), opt_states.data_ptr(),
                                   self._stream())
        if rc != 0:
            _cabi.check(rc)
        return u_opt, opt_states

    def forward_action(self, state: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """First half of a two-stage ``forward_host``: returns the optimal action sequence [T,2] (HOST tensor) as soon as
        the kernel has written it, before the serial optimal rollout (mppi.py:205-213) -- what a control loop needs to step
        the environment (test/test_mppi.py:183-186).  ``wait_states()`` returns the optimal state sequence of the same
        iteration; call it before the next ``forward_host`` / ``forward_action`` (the staging buffer is reused), or skip
        it.  Bit-identical to ``forward_host``'s results."""
        d = self.__dict__
        if d["_slow_host_path"]:
            u_opt, opt_states = self.forward_host(state)
            d["_pending_states"] = opt_states
            if out is not None:
                out.copy_(u_opt)
                return out
            return u_opt
        if out is None:
            out = torch.empty(self._horizon, 2, dtype=torch.float32)
        elif out is not d.get("_host_action_ok"):
            if not (out.dtype == torch.float32 and out.device.type == "cpu" and out.is_contiguous()
                    and out.shape == (self._horizon, 2)):
                raise ValueError("out must be a contiguous fp32 CPU tensor of shape [T,2]")
            d["_host_action_ok"] = out
        if state is not d.get("_host_state_ok"):
            if not (torch.is_tensor(state) and state.dtype == torch.float32 and state.device.type == "cpu"
                    and state.is_contiguous()):
                state = torch.as_tensor(state, dtype=torch.float32).detach().cpu().contiguous()
            assert state.shape == (self._dim_state,)
            d["_host_state_ok"] = state
        self._sync_problem()
        if d["_action_noises"] is not d["_engine_noise"]:
            d["_action_noises"] = d["_engine_noise"]
        rc = d["_forward_action_fn"](d["_handle"], state.data_ptr(), out.data_ptr(), self._stream())
        if rc != 0:
            _cabi.check(rc)
        d["_pending_states"] = None
        return out

    def wait_states(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Second half of ``forward_action``: the optimal state sequence [1,T+1,3] (HOST tensor) of that iteration."""
        d = self.__dict__
        if out is None:
            out = torch.empty(1, self._horizon + 1, 3, dtype=torch.float32)
        elif not (out.dtype == torch.float32 and out.device.type == "cpu" and out.is_contiguous()
                  and out.shape == (1, self._horizon + 1, 3)):
            raise ValueError("out must be a contiguous fp32 CPU tensor of shape [1,T+1,3]")
        pending = d.get("_pending_states")
        if pending is not None:  # (slow host path: forward_action already has them)
            out.copy_(pending)
            return out
        _cabi.check(self._lib.bnv_mppi_wait_states(self._handle, out.data_ptr()))
        return out

    def forward_follow(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """The other ranks' half of a host-driven sharded iteration: while ONE rank (the leader) calls
        ``forward_host(state)``, every other rank calls this once per leader call.  The kernel waits on the device for
        the leader's state (broadcast over NVLink by the leader's kernel), rolls out this rank's shard and takes part
        in the exchange.  Asynchronous; returns this rank's device copies of the (identical) results."""
        u_opt = torch.empty(self._horizon, 2, device=self._device, dtype=torch.float32)
        opt_states = torch.empty(1, self._horizon + 1, 3, device=self._device, dtype=torch.float32)
        self._sync_problem()
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_mppi_forward_follow(self._handle, None, u_opt.data_ptr(), opt_states.data_ptr(),
                                                          self._stream()))
        return u_opt, opt_states

    def get_top_samples(self, num_samples: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Top ``num_samples`` rollouts by weight, descending (mppi.py:221-240).  A solver built with
        ``record_states=False`` keeps no state array: the selected samples are rolled out again (bit-identical rows)."""
        assert num_samples <= self._num_samples
        n_local = min(int(num_samples), self._local_samples)
        states = torch.empty(n_local, self._horizon + 1, 3, device=self._device, dtype=torch.float32)
        weights = torch.empty(n_local, device=self._device, dtype=torch.float32)
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_mppi_top_samples(self._handle, n_local, states.data_ptr(), weights.data_ptr(),
                                                       self._stream()))
        if self._shard.world_size > 1:
            with torch.cuda.device(self._device):
                return merge_top_candidates(states, weights, int(num_samples), self._shard, self._lib, self._handle,
                                            self._stream())
        return states, weights

    # ------------------------------------------------------------------ extras
    def draw_lookup_normals(self, iteration: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(xi [K,2T+1], xi_opt [T]) the engine draws for its stochastic lookups in the given iteration (0-based since
        construction / reset): lets a test replay an in-engine-noise iteration through the oracle."""
        xi = torch.empty(self._local_samples, 2 * self._horizon + 1, device=self._device, dtype=torch.float32)
        xi_opt = torch.empty(self._horizon, device=self._device, dtype=torch.float32)
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_mppi_draw_xi(self._handle, int(iteration), xi.data_ptr(), xi_opt.data_ptr(),
                                                   self._stream()))
        return xi, xi_opt

    def reset(self) -> None:
        """Zero the mean sequence and restart the engine's noise stream (a freshly built solver)."""
        _cabi.check(self._lib.bnv_mppi_reset(self._handle, self._stream()))

    def prelaunch(self, enable: bool = True, timeout_us: int = 2000) -> None:
        """Pre-launch the next iteration's kernel from every ``forward_host`` call (``bnv_mppi_prelaunch``): the kernel is
        resident and polling a host-mapped mailbox when the next state arrives, which takes the launch call and the
        launch latency off the control loop's critical path.  Results are identical to the plain path."""
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_mppi_prelaunch(self._handle, 1 if enable else 0, int(timeout_us)))

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

    def check(self) -> None:
        """Synchronise and raise if an in-kernel wait of an earlier iteration timed out (a peer rank of a sharded
        solver never delivered its softmax partial)."""
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_mppi_check(self._handle, self._stream()))

    @property
    def launch_geometry(self) -> dict:
        """Rollout-kernel launch shape: CTAs, warps per CTA, recorded-state slab split step (0 = one flush), cooperative."""
        out = (C.c_int32 * 4)()
        _cabi.check(self._lib.bnv_mppi_launch_geometry(self._handle, out))
        return {"ctas": out[0], "warps_per_cta": out[1], "rec_split": out[2], "cooperative": bool(out[3])}

    def kernel_timing(self, max_launches: int) -> None:
        """Record CUDA-event pairs around the rollout kernel of the next ``max_launches`` iterations."""
        _cabi.check(self._lib.bnv_mppi_kernel_timing(self._handle, int(max_launches)))

    def kernel_time(self) -> Tuple[float, int]:
        """(summed rollout-kernel milliseconds, launches measured) since kernel_timing(); synchronises."""
        ms, n = C.c_double(), C.c_uint64()
        _cabi.check(self._lib.bnv_mppi_kernel_time(self._handle, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    @property
    def costs(self) -> torch.Tensor:
        """Per-sample costs of the last iteration (mppi.py:186-190; the reference does not keep them)."""
        return self._costs
