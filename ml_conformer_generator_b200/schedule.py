"""Host side of the noise schedule: the gamma table and the per-step scalar coefficients the CUDA step kernel
consumes.  Mirrors reference equivariant_diffusion.py:9-45 (polynomial_schedule / clip_noise_schedule),
:108-134 (PredefinedNoiseSchedule), :224-247 (sigma_and_alpha_t_given_s) and :305-326 (sample_p_zs_given_zt) in the
reference's own float32 torch arithmetic, evaluated once on the host: the scalars are identical for the whole batch."""
from typing import Dict, List

import torch
import torch.nn.functional as F


def gamma_table(timesteps: int, precision: float = 1e-5, power: int = 2) -> torch.Tensor:
    steps = timesteps + 1
    grid = torch.linspace(0, steps, steps)
    a2 = (1 - torch.pow(grid / steps, power)) ** 2
    padded = torch.cat((torch.ones(1), a2), dim=0)
    a2 = torch.cumprod(torch.clip(padded[1:] / padded[:-1], min=0.001, max=1.0), dim=0)
    a2 = (1 - 2 * precision) * a2 + precision
    return (-(torch.log(a2) - torch.log(1 - a2))).float()


def _gamma_at(gamma: torch.Tensor, frac: torch.Tensor) -> torch.Tensor:
    return gamma[torch.round(frac * (gamma.numel() - 1)).long()]


def step_scalars(gamma: torch.Tensor, s: int) -> Dict[str, float]:
    """z_s = c_z * z_t - c_eps * eps + c_sigma * noise   (then COM removal of the x part), t = (s+1)/T."""
    T = gamma.numel() - 1
    s_i = torch.full([1], s, dtype=torch.int64)
    t_f = (s_i + 1.0) / T
    s_f = s_i / T
    g_s, g_t = _gamma_at(gamma, s_f), _gamma_at(gamma, t_f)
    sigma2_ts = 1 - torch.exp(F.softplus(g_s) - F.softplus(g_t))
    alpha_ts = torch.exp(0.5 * (F.logsigmoid(-g_t) - F.logsigmoid(-g_s)))
    sigma_s = torch.sqrt(torch.sigmoid(g_s))
    sigma_t = torch.sqrt(torch.sigmoid(g_t))
    return {
        "t": float(t_f), "s": float(s_f),
        "alpha_ts": float(alpha_ts),
        "c_eps": float(sigma2_ts / alpha_ts / sigma_t),
        "c_sigma": float(torch.sqrt(sigma2_ts) * sigma_s / sigma_t),
        "alpha_s": float(torch.sqrt(torch.sigmoid(-g_s))),
        "sigma_s": float(sigma_s),
    }


def decode_scalars(gamma: torch.Tensor) -> Dict[str, float]:
    """Scalars of sample_p_xh_given_z0 (reference equivariant_diffusion.py:269-277): x = (z0 - sigma0*eps)/alpha0 +
    sigma_x*noise."""
    g0 = gamma[0:1]
    return {
        "sigma_0": float(torch.sqrt(torch.sigmoid(g0))),
        "alpha_0": float(torch.sqrt(torch.sigmoid(-g0))),
        "sigma_x": float(torch.exp(0.5 * g0)),
    }


def forward_level_scalars(gamma: torch.Tensor, level: int) -> Dict[str, float]:
    """alpha/sigma at integer level (merge_fragments' initial forward diffusion, reference :549-558)."""
    T = gamma.numel() - 1
    g = _gamma_at(gamma, torch.full([1], level, dtype=torch.int64) / T)
    return {"alpha": float(torch.sqrt(torch.sigmoid(-g))), "sigma": float(torch.sqrt(torch.sigmoid(g)))}


def all_step_scalars(gamma: torch.Tensor) -> List[Dict[str, float]]:
    T = gamma.numel() - 1
    return [step_scalars(gamma, s) for s in range(T)]
