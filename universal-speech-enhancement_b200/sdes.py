"""SDE classes of the path (host side; scalars and schedule tables only -- tensors stay in the CUDA kernels).

Mirrors /root/reference/src/models/components/sgmse/sdes.py: ``SDERegistry``, ``SDE.discretize`` (:75-92),
``SDE.reverse`` -> ``ReverseSDE`` (:94-175) and ``OUVESDE`` (:182-279).  OUVPSDE is out of scope (not selectable from the
shipped configs).  The tensor methods are written with torch ops so they work on any
device for the non-fused sampler variants; the fused predictor consumes ``step_tables`` instead.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch

from .registry import Registry

SDERegistry = Registry("SDE")


class SDE:
    def __init__(self, N):
        self.N = N

    @property
    def T(self):
        raise NotImplementedError

    def sde(self, x, t, *args):
        raise NotImplementedError

    def discretize(self, x, t, *args):
        """x_{i+1} = x_i + f_i + G_i z_i with dt = 1/N (NOT (T-eps)/N: a reference quirk that is preserved)."""
        dt = 1 / self.N
        drift, diffusion = self.sde(x, t, *args)
        f = drift * dt
        G = diffusion * torch.sqrt(torch.tensor(dt, device=t.device))
        return f, G

    def reverse(self, score_model, probability_flow=False):
        """The reverse-time SDE / probability-flow ODE around ``score_model`` (sdes.py:94-175 of the reference)."""
        return ReverseSDE(self, score_model, probability_flow)

    def copy(self):
        raise NotImplementedError


class ReverseSDE:
    """dx = [f(x, t) - g(t)^2 score(x, t) c] dt + g(t) dw_bar with c = 1 (SDE) or 1/2 and no noise (probability flow).

    Host-side helper for third-party predictors / correctors (the built-in ones run fused in CUDA and never touch it):
    the score comes from the B200 network (one C call), the elementwise part is torch on the device.  Same two entry
    points as the reference's RSDE: ``sde(x, t, y, conditioning=...)`` -> (drift, diffusion) and
    ``discretize(...)`` -> (f, G) with the forward SDE's dt = 1/N.
    """

    def __init__(self, forward_sde, score_model, probability_flow=False):
        self.fwd, self.score_model = forward_sde, score_model
        self.N, self.probability_flow = forward_sde.N, probability_flow

    @property
    def T(self):
        return self.fwd.T

    def _score_term(self, x, t, y, g, conditioning):
        cond = [y] if conditioning is None else conditioning
        score = self.score_model(x, t, score_conditioning=cond, sde_input=y)
        g = g.reshape(g.shape + (1,) * (x.ndim - g.ndim))
        weight = 0.5 if self.probability_flow else 1.0
        return g, g**2 * score * weight

    def sde(self, x, t, y, conditioning=None):
        drift, g = self.fwd.sde(x, t, y)
        g, term = self._score_term(x, t, y, g, conditioning)
        return drift - term, (torch.zeros_like(g) if self.probability_flow else g)

    def discretize(self, x, t, y, conditioning=None):
        f, G = self.fwd.discretize(x, t, y)
        G, term = self._score_term(x, t, y, G, conditioning)
        return f - term, (torch.zeros_like(G) if self.probability_flow else G)


@SDERegistry.register("ouve")
class OUVESDE(SDE):
    """dx = theta (y - x) dt + sigma_min (sigma_max/sigma_min)^t sqrt(2 log(sigma_max/sigma_min)) dw."""

    def __init__(self, theta=1.5, sigma_min=0.05, sigma_max=0.5, N=1000, **ignored_kwargs):
        super().__init__(N)
        self.theta = theta
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.logsig = np.log(self.sigma_max / self.sigma_min)

    def copy(self):
        return OUVESDE(self.theta, self.sigma_min, self.sigma_max, N=self.N)

    @property
    def T(self):
        return 1

    def sde(self, x, t, y):
        drift = self.theta * (y - x)
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        return drift, sigma * np.sqrt(2 * self.logsig)

    def _mean(self, x0, t, y):
        e = torch.exp(-self.theta * t)[:, None, None, None]
        return e * x0 + (1 - e) * y

    def _std(self, t, **kwargs):
        s, th, ls = self.sigma_min, self.theta, self.logsig
        return torch.sqrt((s**2 * torch.exp(-2 * th * t) * (torch.exp(2 * (th + ls) * t) - 1) * ls) / (th + ls))

    def marginal_prob(self, x0, t, y):
        return self._mean(x0, t, y), self._std(t)

    def prior_sampling(self, shape, y):
        if shape != y.shape:
            warnings.warn(f"Target shape {shape} does not match shape of y {y.shape}! Ignoring target shape.")
        std = self._std(torch.ones((y.shape[0],), device=y.device))
        return y + torch.randn_like(y) * std[:, None, None, None]

    def prior_logp(self, z):
        raise NotImplementedError("prior_logp for OU SDE not yet implemented!")

    # ---- what the fused CUDA sampler consumes ----------------------------------------------------
    def step_tables(self, N: int, eps: float, times=None):
        """(t_i, G_i, std(T)) as CPU float32: the bit-exact step schedule, plus the tables of the sampler variants.

        t_i = torch.linspace(T, eps, N) (sampling/__init__.py:63) and G_i = g(t_i) * sqrt(float32(1/N))
        (sdes.py:88-92,216-224) are evaluated with the SAME torch CPU expressions the reference uses, so the
        integer step index i in [0, N) maps to identical float32 bit patterns (tests/test_schedule.py).
        ``times`` replaces the linspace (single update_fn steps); N stays the dt denominator.
        """
        ts = torch.linspace(self.T, eps, N) if times is None else times.detach().to("cpu", torch.float32)
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** ts
        G = sigma * np.sqrt(2 * self.logsig) * torch.sqrt(torch.tensor(1 / N))
        std1 = self._std(torch.ones((1,)))
        return ts, G, float(std1[0])

    def variant_tables(self, ts: torch.Tensor, snr: float):
        """float32 tables of the other registered steps, with the reference's torch expressions:
        g(t_i) (OUVESDE.sde, sdes.py:216-224; Euler-Maruyama) and the annealed-Langevin step size
        2 (snr std(t_i))^2 (correctors.py:84,94)."""
        ts = ts.detach().to("cpu", torch.float32)
        g = self.sigma_min * (self.sigma_max / self.sigma_min) ** ts * np.sqrt(2 * self.logsig)
        ald = (snr * self._std(ts)) ** 2 * 2
        return g.to(torch.float32).contiguous(), ald.to(torch.float32).contiguous()
