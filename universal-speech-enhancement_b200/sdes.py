"""SDE classes of the path (host side; scalars and schedule tables only -- tensors stay in the CUDA kernels).

Mirrors /root/reference/src/models/components/sgmse/sdes.py: ``SDERegistry``, ``SDE.discretize`` (:75-92),
``SDE.reverse`` -> RSDE (:94-175) and ``OUVESDE`` (:182-279).  OUVPSDE is out of scope (not selectable from the
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

    def reverse(oself, score_model, probability_flow=False):
        N, T, sde_fn, discretize_fn = oself.N, oself.T, oself.sde, oself.discretize

        class RSDE(oself.__class__):
            def __init__(self):
                self.N = N
                self.probability_flow = probability_flow

            @property
            def T(self):
                return T

            def _score(self, x, t, *args, **kwargs):
                if kwargs.get("conditioning") is not None:
                    return score_model(x, t, score_conditioning=kwargs["conditioning"], sde_input=args[0])
                return score_model(x, t, *args)

            def sde(self, x, t, *args, **kwargs):
                drift, diffusion = sde_fn(x, t, *args)
                score = self._score(x, t, *args, **kwargs)
                if diffusion.ndim < x.ndim:
                    diffusion = diffusion.view(*diffusion.size(), *((1,) * (x.ndim - diffusion.ndim)))
                total = drift - diffusion**2 * score * (0.5 if self.probability_flow else 1.0)
                return total, (torch.zeros_like(diffusion) if self.probability_flow else diffusion)

            def discretize(self, x, t, *args, **kwargs):
                f, G = discretize_fn(x, t, *args)
                if G.ndim < x.ndim:
                    G = G.view(*G.size(), *((1,) * (x.ndim - G.ndim)))
                rev_f = f - G**2 * self._score(x, t, *args, **kwargs) * (0.5 if self.probability_flow else 1.0)
                return rev_f, (torch.zeros_like(G) if self.probability_flow else G)

        return RSDE()

    def copy(self):
        raise NotImplementedError


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
    def step_tables(self, N: int, eps: float):
        """(t_i, G_i, std(T)) as CPU float32: the bit-exact step schedule.

        t_i = torch.linspace(T, eps, N) (sampling/__init__.py:63) and G_i = g(t_i) * sqrt(float32(1/N))
        (sdes.py:88-92,216-224) are evaluated with the SAME torch CPU expressions the reference uses, so the
        integer step index i in [0, N) maps to identical float32 bit patterns (tests/test_schedule.py).
        """
        ts = torch.linspace(self.T, eps, N)
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** ts
        G = sigma * np.sqrt(2 * self.logsig) * torch.sqrt(torch.tensor(1 / N))
        std1 = self._std(torch.ones((1,)))
        return ts, G, float(std1[0])
