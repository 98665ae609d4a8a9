"""ORACLE (test infrastructure): numpy/torch restatement of ResShift's diffusion schedule and
the ``p_sample`` residual-shift loop.  Every function cites the reference file:line it follows;
pinned against tables and trajectories produced by the imported reference
(``oracle/make_golden.py`` -> ``tests/golden/``) and against the known-answer constants of
SURVEY.md Appendix B.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import numpy as np
import torch


def eta_schedule(steps: int, min_noise_level: float, etas_end: float, kappa: float, power: float) -> np.ndarray:
    """sqrt_etas of the 'exponential' schedule — reference models/gaussian_diffusion.py:45-58."""
    etas_start = min(min_noise_level / kappa, min_noise_level)
    increaser = math.exp(1.0 / (steps - 1) * math.log(etas_end / etas_start))
    base = np.ones([steps]) * increaser
    power_timestep = np.linspace(0, 1, steps, endpoint=True) ** power
    power_timestep *= (steps - 1)
    return np.power(base, power_timestep) * etas_start


def schedule_tables(sqrt_etas: np.ndarray, kappa: float) -> Dict[str, np.ndarray]:
    """Posterior coefficients — reference models/gaussian_diffusion.py:143-161,598-603 (float64)."""
    etas = sqrt_etas ** 2
    etas_prev = np.append(0.0, etas[:-1])
    alpha = etas - etas_prev
    post_var = kappa ** 2 * etas_prev / etas * alpha
    post_var_clipped = np.append(post_var[1], post_var[1:])
    return {
        "sqrt_etas": sqrt_etas,
        "etas": etas,
        "coef1": etas_prev / etas,                       # multiplies x_t
        "coef2": alpha / etas,                           # multiplies pred_xstart
        "log_var": np.log(post_var_clipped),
        "std": np.exp(0.5 * np.log(post_var_clipped)),   # what p_sample multiplies the noise by
        "in_scale": 1.0 / np.sqrt(etas * kappa ** 2 + 1.0),   # _scale_input, latent_flag=True
    }


def p_sample_loop(model: Callable, z_y: torch.Tensor, noises: List[torch.Tensor], tabs: Dict[str, np.ndarray],
                  kappa: float, record: Optional[list] = None) -> torch.Tensor:
    """reference models/gaussian_diffusion.py:421-472 (+ p_sample :332-365, p_mean_variance :234-307,
    prior_sample :517-529) for predict_type 'xstart', normalize_input & latent_flag True, no clipping.

    ``model(x_in, t)`` returns pred_xstart; ``noises`` holds T+1 tensors in draw order (prior first,
    then one per step including the unused one at t == 0, as the reference draws them).
    """
    T = len(tabs["etas"])
    f32 = lambda a, i: torch.tensor(float(np.float32(a[i])))   # _extract_into_tensor casts to fp32 (:102)
    x = z_y + f32(kappa * tabs["sqrt_etas"], T - 1) * noises[0]
    for k, i in enumerate(range(T - 1, -1, -1)):
        t = torch.full((z_y.shape[0],), i, dtype=torch.long, device=z_y.device)
        std_in = torch.sqrt(f32(tabs["etas"], i) * kappa ** 2 + 1)
        pred = model(x / std_in, t).float()
        mean = f32(tabs["coef1"], i) * x + f32(tabs["coef2"], i) * pred
        nonzero = 0.0 if i == 0 else 1.0
        sample = mean + nonzero * torch.exp(0.5 * f32(tabs["log_var"], i)) * noises[k + 1]
        if record is not None:
            record.append({"sample": sample, "pred_xstart": pred, "mean": mean})
        x = sample
    return x
