"""Risk-map inference on the device: ``TraversabilityModel._infer_risk_map``
(src/simulator/problem_formulation/traversability_model.py:28-51), the step immediately before the planner.

``method="closed_form"`` (default) evaluates expected value / VaR / CVaR of the per-cell Normal slip model exactly
(mean + coef * std); ``method="monte_carlo"`` is the reference's estimator (``num_samples`` draws per cell,
``torch.quantile`` and the mean of the tail), on injected ``samples`` [S,G,G] or on the engine's own Philox draws.
"""

from __future__ import annotations

from typing import Optional

import torch

from . import _cabi

_METRICS = {"expected_value": 0, "var": 1, "cvar": 2}


def infer_risk_map(mean: torch.Tensor, std: torch.Tensor, inference_metric: str, confidence_value: Optional[float] = None,
                   *, method: str = "closed_form", num_samples: int = 1000, samples: Optional[torch.Tensor] = None,
                   seed: int = 0, return_samples: bool = False):
    if inference_metric not in _METRICS:  # utils.py:18-24
        raise AssertionError(f"inference_metric must be one of {list(_METRICS)}")
    if inference_metric != "expected_value":
        assert confidence_value is not None and 0.0 <= confidence_value <= 1.0, \
            "confidence_value must be set between 0 and 1 when inference_metric is 'var' or 'cvar'."  # utils.py:27-31
    if method not in ("closed_form", "monte_carlo"):
        raise ValueError("method must be 'closed_form' or 'monte_carlo'")
    dev = mean.device
    if dev.type != "cuda":
        raise RuntimeError("benchnav_b200.risk runs on CUDA (sm_100a) only; there is no CPU fallback")
    lib = _cabi.load()
    mean = mean.detach().to(dev, torch.float32).contiguous()
    std = std.detach().to(dev, torch.float32).contiguous()
    n = mean.numel()
    out = torch.empty_like(mean)
    samples_ptr, drawn = None, None
    if samples is not None:
        samples = samples.detach().to(dev, torch.float32).contiguous()
        if samples.dim() != mean.dim() + 1 or tuple(samples.shape[1:]) != tuple(mean.shape):
            raise ValueError("samples must be [num_samples, *mean.shape]")
        num_samples = int(samples.shape[0])
        samples_ptr = samples.data_ptr()
    elif return_samples and method == "monte_carlo" and inference_metric != "expected_value":
        drawn = torch.empty((int(num_samples),) + tuple(mean.shape), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _cabi.check(lib.bnv_risk_map(_METRICS[inference_metric], float(confidence_value or 0.0),
                                     _cabi.BNV_RISK_MONTE_CARLO if method == "monte_carlo" else _cabi.BNV_RISK_CLOSED_FORM,
                                     mean.data_ptr(), std.data_ptr(), n, samples_ptr, int(num_samples), int(seed),
                                     out.data_ptr(), drawn.data_ptr() if drawn is not None else None,
                                     torch.cuda.current_stream(dev).cuda_stream))
    return (out, drawn) if return_samples else out
