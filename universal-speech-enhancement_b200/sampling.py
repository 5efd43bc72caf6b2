"""Predictor-corrector sampling: the reference's plugin surface over the fused CUDA loop.

Mirrors the API of /root/reference/src/models/components/sgmse/sampling/{__init__,predictors,correctors}.py -- the
``PredictorRegistry`` / ``CorrectorRegistry`` names ("reverse_diffusion", "euler_maruyama", "none"; "langevin",
"ald", "none"), ``update_fn(x, t, y, conditioning=...) -> (x, x_mean)`` and ``get_pc_sampler(...) -> callable``
returning ``(x_result, n_function_evaluations)`` -- but none of the arithmetic lives here:

  * every built-in predictor / corrector is a *descriptor* (``kind``) of a step the C library executes; the whole
    sampler (prior draw + N x [corrector steps, predictor step], each = one NCSN++ evaluation + one fused update
    kernel) is ONE call of ``use_pc_sample_ex`` (include/use_b200.h) with no host synchronisation inside;
  * ``update_fn`` of a built-in class is the same C call restricted to one step starting from the given state, so
    third-party code that drives the classes by hand gets the same kernels;
  * predictors / correctors registered by third parties (no ``kind``) run through ``_host_loop`` below, which only
    sequences their ``update_fn`` calls.

Reference quirks kept or documented:
  * draws are consumed in the reference's order (prior, then per outer step the corrector's draws, then the
    predictor's); ``noise`` makes them explicit for parity tests, otherwise Philox streams keyed by the global clip
    index are used (shard invariant);
  * ``LangevinCorrector`` couples the batch through two batch-mean norms (correctors.py:55-57): under sharding or
    micro-batching the mean is over the local (micro-)batch, exactly as the reference's per-rank / ``minibatch``
    behaviour;
  * ``probability_flow`` never reaches the reverse SDE of a predictor in the reference (predictors.py:14-19 builds
    ``sde.reverse(score_fn)`` without it), so it does not change the PC sampler here either (the C library's
    ``use_sampler_opts.probability_flow`` implements the halved-score / no-noise step for callers that want it);
  * ``euler_maruyama`` cannot run in the reference with this ScoreModel (RSDE.rsde_parts calls
    ``score_model(x, t, conditioning)`` without ``sde_input``, sdes.py:131-134 -> TypeError); here it follows the
    evident intent and is pinned against the reference's own classes driven with an adapter score function
    (oracle/make_golden_variants.py).
"""
from __future__ import annotations

import abc

import torch

from . import _lib, sdes
from .registry import Registry

PredictorRegistry = Registry("Predictor")
CorrectorRegistry = Registry("Corrector")


def _uniform_time(t: torch.Tensor) -> float:
    """update_fn receives vec_t = ones(B) * t_i (sampling/__init__.py:65-66): the fused step takes the scalar."""
    t = t.detach().reshape(-1)
    t0 = float(t[0])
    if t.numel() > 1 and not bool((t == t[0]).all()):
        raise NotImplementedError("the fused sampler steps take one batch-uniform time value")
    return t0


class _Step(abc.ABC):
    """Common part of the built-in predictors / correctors: one fused step through the C library."""

    kind: str = None  # name of the fused step; None for host-defined plugins

    def _fused_step(self, x, t, y, conditioning, **sel):
        fn = getattr(self.score_fn, "_fused_pc_sample", None)
        if fn is None:
            raise NotImplementedError(f"{type(self).__name__} needs a B200 ScoreModel as score_fn")
        if conditioning is not None and len(conditioning) not in (1, 2):
            raise NotImplementedError("fused steps take one or two conditioning spectrograms")
        cond = None if conditioning is None or conditioning[0] is y else conditioning[0]
        cond2 = conditioning[1] if conditioning is not None and len(conditioning) == 2 else None
        return fn(self.sde, y, None, x_init=x, times=torch.tensor([_uniform_time(t)]), want_state=True, cond=cond, cond2=cond2,
                  **sel)


class Predictor(_Step):
    """Base class of the plugin API (predictors.py:11-38)."""

    def __init__(self, sde, score_fn, probability_flow=False):
        # reference quirk (predictors.py:14-19): the flag is stored but the reverse SDE is always built WITHOUT it, so
        # get_pc_sampler(probability_flow=True) samples exactly like probability_flow=False; kept.
        self.sde, self.score_fn, self.probability_flow = sde, score_fn, probability_flow
        self.rsde = sde.reverse(score_fn)

    @abc.abstractmethod
    def update_fn(self, x, t, *args, **kwargs):
        """One predictor update: returns (x, x_mean)."""


class Corrector(_Step):
    """Base class of the plugin API (correctors.py:11-34)."""

    def __init__(self, sde, score_fn, snr, n_steps):
        self.sde, self.score_fn, self.snr, self.n_steps = sde, score_fn, snr, n_steps
        self.rsde = sde.reverse(score_fn)

    @abc.abstractmethod
    def update_fn(self, x, t, *args, **kwargs):
        """One corrector update (n_steps inner steps): returns (x, x_mean)."""


class _FusedPredictor(Predictor):
    def update_fn(self, x, t, y, conditioning=None, **_):
        return self._fused_step(x, t, y, conditioning, predictor=self.kind, corrector="none")


class _FusedCorrector(Corrector):
    def update_fn(self, x, t, y, conditioning=None, **_):
        return self._fused_step(x, t, y, conditioning, predictor="none", corrector=self.kind,
                                corrector_steps=self.n_steps, snr=self.snr, denoise=2)  # 2: the corrector's own x_mean


@PredictorRegistry.register("euler_maruyama")
class EulerMaruyamaPredictor(_FusedPredictor):
    kind = "euler_maruyama"


@PredictorRegistry.register("reverse_diffusion")
class ReverseDiffusionPredictor(_FusedPredictor):
    kind = "reverse_diffusion"


@PredictorRegistry.register("none")
class NonePredictor(Predictor):
    kind = "none"

    def __init__(self, *args, **kwargs):
        self.probability_flow = False

    def update_fn(self, x, t, *args, **kwargs):
        return x, x


@CorrectorRegistry.register(name="langevin")
class LangevinCorrector(_FusedCorrector):
    kind = "langevin"


@CorrectorRegistry.register(name="ald")
class AnnealedLangevinDynamics(_FusedCorrector):
    kind = "ald"

    def __init__(self, sde, score_fn, snr, n_steps):
        if not isinstance(sde, sdes.OUVESDE):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
        super().__init__(sde, score_fn, snr, n_steps)


@CorrectorRegistry.register(name="none")
class NoneCorrector(Corrector):
    kind = "none"

    def __init__(self, *args, **kwargs):
        self.snr, self.n_steps = 0, 0

    def update_fn(self, x, t, *args, **kwargs):
        return x, x


_BUILTIN = {EulerMaruyamaPredictor, ReverseDiffusionPredictor, NonePredictor, LangevinCorrector,
            AnnealedLangevinDynamics, NoneCorrector}  # exact classes: a subclass overriding update_fn is a host plugin


def _fusable(predictor_cls, corrector_cls, sde, score_fn, conditioning, y) -> bool:
    return (predictor_cls in _BUILTIN and corrector_cls in _BUILTIN and isinstance(sde, sdes.OUVESDE) and hasattr(score_fn, "_fused_pc_sample") and sde.N >= 1
            and conditioning is not None and len(conditioning) in (1, 2))


def _host_loop(predictor, corrector, sde, y, eps, denoise, conditioning):
    """Sequencing only, for predictor / corrector classes defined outside this package: prior draw, then the
    corrector-then-predictor order of the reference's pc_sampler over the float32 schedule linspace(T, eps, N)."""
    state = sde.prior_sampling(y.shape, y)
    mean = state
    ones = torch.ones(y.shape[0], device=y.device)
    for t_i in sde.step_tables(sde.N, eps)[0].to(y.device):
        for stage in (corrector, predictor):
            state, mean = stage.update_fn(state, ones * t_i, y, conditioning=conditioning)
    return (mean if denoise and sde.N else state), sde.N * (corrector.n_steps + 1)


def get_pc_sampler(predictor_name, corrector_name, sde, score_fn, y, denoise=True, eps=3e-2, snr=0.1,
                   corrector_steps=1, probability_flow: bool = False, conditioning=None, intermediate=False,
                   noise=None, seed=None, clip0=0, trace=None, job_clips=None, **kwargs):
    """Create a PC sampler (sampling/__init__.py:23-73).  Additions for reproducible / shard-invariant sampling and
    parity tests: ``noise`` (complex [1 + N * draws_per_step, *y.shape], the explicit normal draws), ``seed`` /
    ``clip0`` (Philox streams), ``trace`` (complex [N, *y.shape] device tensor receiving xt_mean of every step),
    ``job_clips`` (clips of the whole job when ``y`` is one shard / minibatch of it: selects the latency or the
    throughput kernels for the whole job, so that its pieces stay bit-identical to the unsplit job)."""
    predictor_cls = PredictorRegistry.get_by_name(predictor_name)
    corrector_cls = CorrectorRegistry.get_by_name(corrector_name)

    if _fusable(predictor_cls, corrector_cls, sde, score_fn, conditioning, y):
        n_corr = 0 if corrector_cls.kind == "none" else corrector_steps

        def fused_sampler():
            out = score_fn._fused_pc_sample(sde, y, eps, predictor=predictor_cls.kind, corrector=corrector_cls.kind,
                                            corrector_steps=corrector_steps, snr=snr, denoise=denoise,
                                            cond=None if conditioning[0] is y else conditioning[0],
                                            cond2=conditioning[1] if len(conditioning) == 2 else None, noise=noise, seed=seed, clip0=clip0, trace=trace,
                                            job_clips=job_clips)
            return out, sde.N * (n_corr + 1)

        return fused_sampler

    if noise is not None or trace is not None:
        raise NotImplementedError("explicit noise / trace need the fused sampler (built-in predictor and corrector)")
    predictor = predictor_cls(sde, score_fn, probability_flow=probability_flow)
    corrector = corrector_cls(sde, score_fn, snr=snr, n_steps=corrector_steps)

    def pc_sampler():
        with torch.no_grad():
            return _host_loop(predictor, corrector, sde, y, eps, denoise, conditioning)

    return pc_sampler


def get_ode_sampler(sde, score_fn, y, inverse_scaler=None, denoise=True, rtol=1e-5, atol=1e-5, method="RK45", eps=3e-2,
                    device=None, conditioning=None, noise=None, **kwargs):
    """Probability-flow ODE sampler with a black-box solver (sampling/__init__.py:76-159): scipy's ``solve_ivp``
    integrates dx/dt = theta (y - x) - g(t)^2 score(x, t) / 2 from T to eps on the HOST (the reference's design: the state
    crosses the PCIe bus as a flattened complex numpy vector at every function evaluation), each right-hand side is ONE
    C call (``use_reverse_drift``: network + fused drift kernel); ``denoise`` adds the noise-free reverse-diffusion step
    at t = eps.  In the reference this sampler cannot run with its own ``ScoreModel`` (``rsde.sde(x, t, y)`` calls
    ``forward()`` without ``sde_input`` -> TypeError, and ``conditioning`` is forwarded into ``solve_ivp``); here it
    follows the evident intent and is pinned against the reference's own ``get_ode_sampler`` driven with an adapter score
    function (oracle/make_golden_variants.py).  ``noise``: complex [1, *y.shape], the explicit prior draw."""
    from scipy import integrate

    fn = getattr(score_fn, "_reverse_drift", None)
    if fn is None or not isinstance(sde, sdes.OUVESDE):
        raise NotImplementedError("get_ode_sampler needs a B200 ScoreModel as score_fn and the OUVE SDE")
    predictor = ReverseDiffusionPredictor(sde, score_fn)
    if conditioning is None:
        conditioning = [y]
    dev = y.device

    def ode_sampler(z=None, **_):
        with torch.no_grad():
            std1 = sde.step_tables(sde.N, eps)[2]
            x = y + noise[0].to(dev) * std1 if noise is not None else sde.prior_sampling(y.shape, y)

            def ode_func(t, x_flat):
                xt = torch.from_numpy(x_flat.reshape(tuple(y.shape))).to(dev).type(torch.complex64)
                drift = fn(sde, xt, y, float(t), conditioning)
                return drift.detach().cpu().numpy().reshape((-1,))

            solution = integrate.solve_ivp(ode_func, (sde.T, eps), x.detach().cpu().numpy().reshape((-1,)), rtol=rtol,
                                           atol=atol, method=method)
            nfe = solution.nfev
            x = torch.tensor(solution.y[:, -1]).reshape(y.shape).to(dev).type(torch.complex64)
            if denoise:  # one predictor step without noise at t = eps
                _, x = predictor.update_fn(x, torch.ones(y.shape[0], device=dev) * eps, y, conditioning=conditioning)
            if inverse_scaler is not None:
                x = inverse_scaler(x)
            return x, nfe

    return ode_sampler
