"""Predictor-corrector sampling (host side).

Mirrors /root/reference/src/models/components/sgmse/sampling/{__init__,predictors,correctors}.py: the
``PredictorRegistry`` / ``CorrectorRegistry`` plugin surface, ``Predictor.update_fn(x, t, *args)`` and
``get_pc_sampler(...) -> callable`` returning ``(x_result, n_function_evaluations)``.

Two execution routes, both with the score network in CUDA (libuse_b200.so):
  * fused: predictor "reverse_diffusion" + corrector "none" + OUVESDE + a B200 ScoreModel as ``score_fn`` --
    the configuration src/predict.py runs (model_wrapper.py:305-314).  The whole N-step loop is ONE C call
    (``use_pc_sample``): prior draw, N x (network + fused drift/diffusion/noise-inject step).
  * generic: any other registered predictor / corrector runs the reference's loop on the host, one C call
    (``use_score_forward``) per score evaluation, elementwise SDE arithmetic as torch ops on the device.
"""
from __future__ import annotations

import abc

import numpy as np
import torch

from . import sdes
from .registry import Registry

PredictorRegistry = Registry("Predictor")
CorrectorRegistry = Registry("Corrector")


class Predictor(abc.ABC):
    def __init__(self, sde, score_fn, probability_flow=False):
        self.sde = sde
        self.rsde = sde.reverse(score_fn)
        self.score_fn = score_fn
        self.probability_flow = probability_flow

    @abc.abstractmethod
    def update_fn(self, x, t, *args):
        """One predictor update: returns (x, x_mean)."""


@PredictorRegistry.register("euler_maruyama")
class EulerMaruyamaPredictor(Predictor):
    def update_fn(self, x, t, *args, **kwargs):
        dt = -1.0 / self.rsde.N
        z = torch.randn_like(x)
        f, g = self.rsde.sde(x, t, *args, **kwargs)
        x_mean = x + f * dt
        if g.ndim < x.ndim:
            g = g.view(*g.size(), *((1,) * (x.ndim - g.ndim)))
        return x_mean + g * np.sqrt(-dt) * z, x_mean


@PredictorRegistry.register("reverse_diffusion")
class ReverseDiffusionPredictor(Predictor):
    def update_fn(self, x, t, *args, **kwargs):
        f, g = self.rsde.discretize(x, t, *args, **kwargs)
        z = torch.randn_like(x)
        x_mean = x - f
        if g.ndim < x.ndim:
            g = g.view(*g.size(), *((1,) * (x.ndim - g.ndim)))
        return x_mean + g * z, x_mean


@PredictorRegistry.register("none")
class NonePredictor(Predictor):
    def __init__(self, *args, **kwargs):
        pass

    def update_fn(self, x, t, *args, **kwargs):
        return x, x


class Corrector(abc.ABC):
    def __init__(self, sde, score_fn, snr, n_steps):
        self.rsde = sde.reverse(score_fn)
        self.score_fn = score_fn
        self.snr = snr
        self.n_steps = n_steps

    @abc.abstractmethod
    def update_fn(self, x, t, *args):
        """One corrector update: returns (x, x_mean)."""

    def _grad(self, x, t, *args, **kwargs):
        if kwargs.get("conditioning") is not None:
            return self.score_fn(x, t, score_conditioning=kwargs["conditioning"], sde_input=args[0])
        return self.score_fn(x, t, *args)


@CorrectorRegistry.register(name="langevin")
class LangevinCorrector(Corrector):
    def update_fn(self, x, t, *args, **kwargs):
        x_mean = x
        for _ in range(self.n_steps):
            grad = self._grad(x, t, *args, **kwargs)
            noise = torch.randn_like(x)
            grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
            noise_norm = torch.norm(noise.reshape(noise.shape[0], -1), dim=-1).mean()
            step_size = ((self.snr * noise_norm / grad_norm) ** 2 * 2).unsqueeze(0)
            step_size = step_size.view(*step_size.size(), *((1,) * (x.ndim - step_size.ndim)))
            x_mean = x + step_size * grad
            x = x_mean + noise * torch.sqrt(step_size * 2)
        return x, x_mean


@CorrectorRegistry.register(name="ald")
class AnnealedLangevinDynamics(Corrector):
    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        if not isinstance(sde, sdes.OUVESDE):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
        self.sde = sde

    def update_fn(self, x, t, *args, **kwargs):
        std = self.sde.marginal_prob(x, t, *args)[1]
        x_mean = x
        for _ in range(self.n_steps):
            grad = self._grad(x, t, *args, **kwargs)
            noise = torch.randn_like(x)
            step_size = (self.snr * std) ** 2 * 2
            step_size = step_size.view(*step_size.size(), *((1,) * (x.ndim - step_size.ndim)))
            x_mean = x + step_size * grad
            x = x_mean + noise * torch.sqrt(step_size * 2)
        return x, x_mean


@CorrectorRegistry.register(name="none")
class NoneCorrector(Corrector):
    def __init__(self, *args, **kwargs):
        self.snr = 0
        self.n_steps = 0

    def update_fn(self, x, t, *args, **kwargs):
        return x, x


def _fusable(predictor_name, corrector_name, sde, score_fn, probability_flow, conditioning, y) -> bool:
    return (predictor_name == "reverse_diffusion" and corrector_name == "none" and isinstance(sde, sdes.OUVESDE)
            and not probability_flow and hasattr(score_fn, "_fused_pc_sample") and sde.N >= 1
            and conditioning is not None and len(conditioning) == 1 and conditioning[0] is y)


def get_pc_sampler(predictor_name, corrector_name, sde, score_fn, y, denoise=True, eps=3e-2, snr=0.1,
                   corrector_steps=1, probability_flow: bool = False, conditioning=None, intermediate=False,
                   noise=None, seed=None, clip0=0, **kwargs):
    """Create a PC sampler (sampling/__init__.py:23-73).  ``noise`` (complex [N+1, *y.shape], explicit draws) /
    ``seed`` / ``clip0`` are additions for reproducible and shard-invariant sampling on the fused route."""
    predictor_cls = PredictorRegistry.get_by_name(predictor_name)
    corrector_cls = CorrectorRegistry.get_by_name(corrector_name)

    if denoise and _fusable(predictor_name, corrector_name, sde, score_fn, probability_flow, conditioning, y):
        def fused_sampler():
            return score_fn._fused_pc_sample(sde, y, eps, noise=noise, seed=seed, clip0=clip0), sde.N

        return fused_sampler

    predictor = predictor_cls(sde, score_fn, probability_flow=probability_flow)
    corrector = corrector_cls(sde, score_fn, snr=snr, n_steps=corrector_steps)

    def pc_sampler():
        with torch.no_grad():
            xt = sde.prior_sampling(y.shape, y).to(y.device)
            timesteps = torch.linspace(sde.T, eps, sde.N).to(y.device)
            xt_mean = xt
            for i in range(sde.N):
                vec_t = torch.ones(y.shape[0], device=y.device) * timesteps[i]
                xt, xt_mean = corrector.update_fn(xt, vec_t, y, conditioning=conditioning)
                xt, xt_mean = predictor.update_fn(xt, vec_t, y, conditioning=conditioning)
            x_result = xt_mean if (denoise and sde.N) else xt
            return x_result, sde.N * (corrector.n_steps + 1)

    return pc_sampler
