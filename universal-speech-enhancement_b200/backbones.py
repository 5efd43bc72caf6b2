"""NCSN++ backbone: reference parameter layout on the host, B200 engine underneath.

Drop-in for /root/reference/src/models/components/sgmse/backbones/ncsnpp.py: same registry names
("ncsnpp", "ncsnpplarge"), same constructor kwargs (unsupported values raise), same ``state_dict`` keys
(``all_modules.{i}.{GroupNorm_0,Conv_0,Dense_0,GroupNorm_1,Conv_1,Conv_2,NIN_k.W/b,W,weight,bias}`` and
``output_layer.*``) so Lightning checkpoints of the reference load with ``strict=True``, and the same
``forward(x: complex [B,2,F,T], time_cond: [B]) -> complex [B,1,F,T]`` (ncsnpp.py:324-501).

The parameters live as ordinary ``nn.Parameter``s; the forward pass does NOT run through torch: the
weights are packed once (OIHW fp32 -> [tap][O][I] bf16 / TF32-rounded fp32) and every layer executes in
libuse_b200.so (tcgen05 implicit-GEMM convolutions and fused bandwidth kernels).  No fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .registry import Registry

BackboneRegistry = Registry("Backbone")


class _Bag(nn.Module):
    """Parameter container; forward lives in CUDA."""


def _uniform_fan_avg(shape, scale, gen=None):
    """DDPM 'default_init': variance_scaling(scale, fan_avg, uniform) (layers.py:66-103); scale 0 -> 1e-10."""
    scale = 1e-10 if scale == 0 else scale
    rf = int(np.prod(shape)) / shape[0] / shape[1]
    fan_in, fan_out = shape[1] * rf, shape[0] * rf
    var = scale / ((fan_in + fan_out) / 2)
    return (torch.rand(*shape, generator=gen) * 2.0 - 1.0) * math.sqrt(3 * var)


def _conv_bag(cout, cin, k, init_scale=1.0):
    b = _Bag()
    b.weight = nn.Parameter(_uniform_fan_avg((cout, cin, k, k), init_scale))
    b.bias = nn.Parameter(torch.zeros(cout))
    return b


def _linear_bag(cout, cin):
    b = _Bag()
    b.weight = nn.Parameter(_uniform_fan_avg((cout, cin), 1.0))
    b.bias = nn.Parameter(torch.zeros(cout))
    return b


def _gn_bag(c):
    b = _Bag()
    b.weight = nn.Parameter(torch.ones(c))
    b.bias = nn.Parameter(torch.zeros(c))
    return b


def _nin_bag(c, init_scale=0.1):
    b = _Bag()
    b.W = nn.Parameter(_uniform_fan_avg((c, c), init_scale))
    b.b = nn.Parameter(torch.zeros(c))
    return b


def module_plan(nf: int, ch_mult: Tuple[int, ...], num_res_blocks: int, input_channels: int,
                conditional: bool = True) -> List[dict]:
    """Construction order of ``all_modules`` (ncsnpp.py:186-316) for the 'biggan' / 'output_skip' /
    'input_skip' / 'sum' / fir configuration, the only one the shipped configs use.  Without the noise
    conditioning (discriminative=True) the two time-embedding Linear layers are absent (ncsnpp.py:196-202)."""
    nres = len(ch_mult)
    plan = [dict(kind="gfp")]
    if conditional:
        plan += [dict(kind="linear", cin=2 * nf, cout=4 * nf), dict(kind="linear", cin=4 * nf, cout=4 * nf)]
    plan.append(dict(kind="conv3", cin=input_channels, cout=nf))
    hs_c, in_ch = [nf], nf
    for lvl in range(nres):
        for _ in range(num_res_blocks):
            out_ch = nf * ch_mult[lvl]
            plan.append(dict(kind="rb", cin=in_ch, cout=out_ch, up=False, down=False))
            in_ch = out_ch
            hs_c.append(in_ch)
        if lvl != nres - 1:
            plan.append(dict(kind="rb", cin=in_ch, cout=in_ch, up=False, down=True))
            plan.append(dict(kind="combine", cin=input_channels, cout=in_ch))
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    plan += [dict(kind="rb", cin=in_ch, cout=in_ch, up=False, down=False), dict(kind="attn", c=in_ch),
             dict(kind="rb", cin=in_ch, cout=in_ch, up=False, down=False)]
    for lvl in reversed(range(nres)):
        for _ in range(num_res_blocks + 1):
            out_ch = nf * ch_mult[lvl]
            plan.append(dict(kind="rb", cin=in_ch + hs_c.pop(), cout=out_ch, up=False, down=False))
            in_ch = out_ch
        plan.append(dict(kind="gn", c=in_ch))
        plan.append(dict(kind="conv3", cin=in_ch, cout=input_channels, init_scale=0.0))
        if lvl != 0:
            plan.append(dict(kind="rb", cin=in_ch, cout=in_ch, up=True, down=False))
    assert not hs_c
    return plan


@BackboneRegistry.register("ncsnpp")
class NCSNpp(nn.Module):
    """NCSN++ score network executed by the B200 engine."""

    _SUPPORTED = dict(scale_by_sigma=True, nonlinearity="swish", resamp_with_conv=True, conditional=True, fir=True,
                      fir_kernel=[1, 3, 3, 1], skip_rescale=True, resblock_type="biggan", progressive="output_skip",
                      progressive_input="input_skip", progressive_combine="sum", embedding_type="fourier",
                      spatial_channels=1, dropout=0.0, centered=False)

    def __init__(self, nf=128, ch_mult=(1, 2, 2, 2), num_res_blocks=1, attn_resolutions=(0,), init_scale=0.0,
                 fourier_scale=16, image_size=256, input_channels=4, discriminative=False, compute_dtype="fp32", **kwargs):
        super().__init__()
        self.discriminative = bool(discriminative)
        if self.discriminative:
            # ncsnpp.py:88-94: no noise conditioning, no 1/t scaling, input = [Re y, Im y]
            kwargs.pop("conditional", None)
            kwargs.pop("scale_by_sigma", None)
            input_channels = 2
        for k, v in kwargs.items():
            if k not in self._SUPPORTED:
                raise TypeError(f"NCSNpp: unknown argument {k!r}")
            want = self._SUPPORTED[k]
            if (list(v) if isinstance(v, (list, tuple)) else v) != want:
                raise NotImplementedError(f"NCSNpp(B200): {k}={v!r} is not supported on this path (only {want!r})")
        if tuple(attn_resolutions) != (0,):
            raise NotImplementedError("NCSNpp(B200): only attn_resolutions=(0,) (bottleneck attention) is supported")
        if input_channels not in ((2,) if self.discriminative else (4, 6)):
            raise NotImplementedError("NCSNpp(B200): 4 input channels (one conditioning spectrogram), 6 (condition='both') "
                                      "or the discriminative 2-channel generator are supported")
        self.conditional = not self.discriminative
        self.scale_by_sigma = not self.discriminative
        self.nf, self.ch_mult, self.num_res_blocks = nf, tuple(ch_mult), num_res_blocks
        self.input_channels = input_channels
        self.num_resolutions = len(self.ch_mult)
        self.compute_dtype = compute_dtype
        self.plan = module_plan(nf, self.ch_mult, num_res_blocks, input_channels, self.conditional)

        self.output_layer = nn.Conv2d(input_channels, 2, 1)  # parameters only; never called
        mods = []
        for m in self.plan:
            k = m["kind"]
            if k == "gfp":
                b = _Bag()
                b.W = nn.Parameter(torch.randn(nf) * fourier_scale, requires_grad=False)
            elif k == "linear":
                b = _linear_bag(m["cout"], m["cin"])
            elif k == "conv3":
                b = _conv_bag(m["cout"], m["cin"], 3, m.get("init_scale", 1.0))
            elif k == "gn":
                b = _gn_bag(m["c"])
            elif k == "combine":
                b = _Bag()
                b.Conv_0 = _conv_bag(m["cout"], m["cin"], 1)
            elif k == "attn":
                b = _Bag()
                b.GroupNorm_0 = _gn_bag(m["c"])
                for j in range(4):
                    setattr(b, f"NIN_{j}", _nin_bag(m["c"], 0.1 if j < 3 else init_scale))
            elif k == "rb":
                b = _Bag()
                b.GroupNorm_0 = _gn_bag(m["cin"])
                b.Conv_0 = _conv_bag(m["cout"], m["cin"], 3)
                b.Dense_0 = _linear_bag(m["cout"], 4 * nf)
                b.GroupNorm_1 = _gn_bag(m["cout"])
                b.Conv_1 = _conv_bag(m["cout"], m["cout"], 3, init_scale)
                if m["cin"] != m["cout"] or m["up"] or m["down"]:
                    b.Conv_2 = _conv_bag(m["cout"], m["cin"], 1)
            else:
                raise AssertionError(k)
            mods.append(b)
        self.all_modules = nn.ModuleList(mods)

        self._engines: Dict[Tuple[int, int], "_Engine"] = {}
        self._spec = None    # STFT parameters / theta handed down by ScoreModel for the engine config
        self._theta = 1.5
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_engine())

    @staticmethod
    def add_argparse_args(parser):
        return parser

    # ---- engine management -----------------------------------------------------------------------
    def invalidate_engine(self):
        """Forget packed weights (call after mutating parameters in place; load_state_dict does it itself)."""
        self._engines = {}

    def engine(self, device: torch.device, dtype=None) -> "_Engine":
        code = _lib.dtype_code(self.compute_dtype if dtype is None else dtype)
        key = (device.index if device.index is not None else torch.cuda.current_device(), code)
        eng = self._engines.get(key)
        if eng is None:
            eng = _Engine(self, torch.device("cuda", key[0]), code, spec=self._spec, theta=self._theta)
            self._engines[key] = eng
        return eng

    def gfp_features(self, t: torch.Tensor) -> torch.Tensor:
        """GaussianFourierProjection(log t) on the HOST with the reference's torch expressions
        (layerspp.py:37-39, ncsnpp.py:352): the arguments 2 pi W log t reach ~1e3 rad, so bit-identical
        features need the identical float32 evaluation order, and t is a [B] (or [N]) vector."""
        t = t.detach().to("cpu", torch.float32)
        W = self.all_modules[0].W.detach().to("cpu", torch.float32)
        x_proj = torch.log(t)[:, None] * W[None, :] * 2 * np.pi
        return torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1).contiguous()

    def forward(self, x: torch.Tensor, time_cond: torch.Tensor = None) -> torch.Tensor:
        """x: complex64 [B, 2, F, T] = cat[x_t, Y] (or [B, 3, F, T] = cat[x_t, Y, Y2], 6 input channels) on a CUDA device
        (discriminative: [B, 1, F, T], no time); returns complex64 [B, 1, F, T]."""
        if not x.is_cuda:
            raise RuntimeError("NCSNpp(B200) runs on CUDA tensors only; there is no CPU path")
        if self.discriminative:
            return self.engine(x.device).net(x[:, 0].contiguous(), None, None).unsqueeze(1)
        xt, Y = x[:, 0].contiguous(), x[:, 1].contiguous()
        if self.input_channels == 6:
            return -self.engine(x.device).score(xt, Y, time_cond, Y2=x[:, 2].contiguous()).unsqueeze(1)
        return self.engine(x.device).net(xt, Y, time_cond).unsqueeze(1)


@BackboneRegistry.register("ncsnpplarge")
class NCSNppLarge(NCSNpp):
    """The ~65 M parameter configuration of configs/model/SGMSE_Large.yaml (ncsnpp.py:504-518)."""

    def __init__(self, **kwargs):
        super().__init__(nf=128, ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2, attn_resolutions=(0,), **kwargs)


class _Engine:
    """One use_engine (weights packed for one device and one compute dtype) plus its cached workspace."""

    def __init__(self, net: NCSNpp, device: torch.device, dtype_code: int, spec=None, theta: float = 1.5):
        self.L = _lib.lib()
        self.net_module = net
        self.device = device
        self.dtype_code = dtype_code
        cfg = _lib.UseConfig()
        cfg.nf, cfg.num_levels, cfg.num_res_blocks = net.nf, len(net.ch_mult), net.num_res_blocks
        for i, m in enumerate(net.ch_mult):
            cfg.ch_mult[i] = m
        cfg.input_channels, cfg.act_dtype = net.input_channels, dtype_code
        sp = spec or {}
        cfg.n_fft, cfg.hop = sp.get("n_fft", 1022), sp.get("hop_length", 160)
        cfg.spec_factor, cfg.spec_abs_exponent = sp.get("spec_factor", 0.15), sp.get("spec_abs_exponent", 0.5)
        cfg.theta = theta
        cfg.conditional, cfg.scale_by_sigma = int(net.conditional), int(net.scale_by_sigma)
        self.cfg = cfg
        with torch.cuda.device(device):
            self.h = self.L.use_engine_create(C.byref(cfg))
            if not self.h:
                raise RuntimeError("use_engine_create failed: " + self.L.use_last_error().decode())
            # two half-batches on two streams (HBM-bound kernels of one overlap the convolutions of the other)
            _lib.check(self.L.use_engine_set_option(self.h, b"overlap_groups", int(os.environ.get("USE_B200_OVERLAP", "2"))),
                       "use_engine_set_option")
            for name, p in net.state_dict().items():
                w = p.detach().to("cpu", torch.float32).contiguous()
                shape = (C.c_int64 * max(w.dim(), 1))(*w.shape)
                _lib.check(self.L.use_engine_set_weight(self.h, name.encode(), w.data_ptr(), shape, w.dim()), name)
            nbytes = C.c_size_t()
            _lib.check(self.L.use_engine_pack(self.h, C.byref(nbytes)), "use_engine_pack")
            self.weights = torch.empty(nbytes.value, dtype=torch.uint8, device=device)
            _lib.check(self.L.use_engine_upload(self.h, self.weights.data_ptr(), nbytes.value, _lib.stream_ptr()),
                       "use_engine_upload")
        self._ws = None
        self._ws_key = None

    def set_option(self, key: str, value: int) -> None:
        """Engine A/B switches (`fuse_gn`, `fuse_head`, `use_graphs`, `overlap_groups`): results are bit-identical either way; the
        workspace plan may change, so the cached workspace is dropped."""
        _lib.check(self.L.use_engine_set_option(self.h, key.encode(), int(value)), "use_engine_set_option")
        if key != "ksplit":  # (the latency mode is part of the program key and does not change the workspace plan)
            self._ws = None
            self._ws_key = None

    def latency_mode(self, job_clips):
        """Pin the engine's latency mode (split-K clusters at the low-resolution levels, include/use_b200.h "ksplit") from
        the size of the WHOLE job for the calls that follow: a job that is cut into micro-batches or shards must run every
        piece in the same mode, otherwise the pieces would differ from the unsplit job in the last bits.  ``None`` = back
        to auto (decided per call: at most two clips).  USE_B200_KSPLIT=0|1 overrides."""
        if os.environ.get("USE_B200_KSPLIT") in ("0", "1"):
            return
        self.set_option("ksplit", 2 if job_clips is None else (1 if int(job_clips) <= 2 else 0))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.use_engine_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def workspace(self, B: int, F: int, T: int) -> torch.Tensor:
        key = (B, F, T)
        if self._ws_key != key:
            n = C.c_size_t()
            _lib.check(self.L.use_engine_workspace_bytes(self.h, B, F, T, C.byref(n)), "use_engine_workspace_bytes")
            self._ws = None  # release before allocating the next one
            self._ws = torch.empty(n.value, dtype=torch.uint8, device=self.device)
            self._ws_key = key
        return self._ws

    def score(self, x: torch.Tensor, Y: torch.Tensor, t: torch.Tensor, Y2: torch.Tensor = None) -> torch.Tensor:
        """-net(cat[x, Y (, Y2)], t) for complex64 [B, F, T] CUDA tensors (Y2: the 6-channel network's second conditioning)."""
        assert x.dtype == torch.complex64 and Y.dtype == torch.complex64 and x.shape == Y.shape and x.dim() == 3
        B, F, T = x.shape
        x, Y = x.contiguous(), Y.contiguous()
        t_host = t.detach().to("cpu", torch.float32).contiguous()
        gfp = self.net_module.gfp_features(t_host)
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            ws = self.workspace(B, F, T)
            if Y2 is not None:
                Y2 = Y2.to(torch.complex64).contiguous()
                assert Y2.shape == x.shape
                _lib.check(self.L.use_score_forward2(self.h, B, F, T, x.data_ptr(), Y.data_ptr(), Y2.data_ptr(),
                                                     t_host.data_ptr(), gfp.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                                     ws.numel(), _lib.stream_ptr()), "use_score_forward2")
            else:
                _lib.check(self.L.use_score_forward(self.h, B, F, T, x.data_ptr(), Y.data_ptr(), t_host.data_ptr(),
                                                    gfp.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                    _lib.stream_ptr()), "use_score_forward")
        return out

    def net(self, x: torch.Tensor, Y, t) -> torch.Tensor:
        """+net(...) = NCSNpp.forward: complex64 [B, F, T] in and out.  Discriminative networks take Y = t = None."""
        assert x.dtype == torch.complex64 and x.dim() == 3
        B, F, T = x.shape
        x = x.contiguous()
        yp = tp = gp = None
        keep = []
        if Y is not None:
            Y = Y.contiguous()
            yp = Y.data_ptr()
        if t is not None:
            t_host = t.detach().to("cpu", torch.float32).contiguous()
            gfp = self.net_gfp(t_host)
            keep += [t_host, gfp]
            tp, gp = t_host.data_ptr(), gfp.data_ptr()
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            ws = self.workspace(B, F, T)
            _lib.check(self.L.use_net_forward(self.h, B, F, T, x.data_ptr(), yp, tp, gp, out.data_ptr(), ws.data_ptr(),
                                              ws.numel(), _lib.stream_ptr()), "use_net_forward")
        return out

    def reverse_drift(self, x, sde_y, t: float, g: float, cond=None, cond2=None, probability_flow=True):
        """use_reverse_drift: theta (y - x) - g^2 score c at the batch-uniform time t; complex64 [B, F, T]."""
        assert x.dtype == torch.complex64 and x.dim() == 3 and x.shape == sde_y.shape
        B, F, T = x.shape
        x, sde_y = x.contiguous(), sde_y.contiguous()
        t_host = torch.full((B,), float(t), dtype=torch.float32)
        gfp = self.net_module.gfp_features(t_host)
        keep = [c.to(torch.complex64).contiguous() if c is not None else None for c in (cond, cond2)]
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            ws = self.workspace(B, F, T)
            _lib.check(self.L.use_reverse_drift(self.h, B, F, T, x.data_ptr(), sde_y.data_ptr(),
                                                keep[0].data_ptr() if keep[0] is not None else None,
                                                keep[1].data_ptr() if keep[1] is not None else None, t_host.data_ptr(),
                                                gfp.data_ptr(), float(g), int(bool(probability_flow)), out.data_ptr(),
                                                ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "use_reverse_drift")
        return out

    def train_forward(self, X0, Y, t, coef, noise=None, seed=0, clip0=0, mae=False):
        """use_train_forward: returns (loss [1 + B] float32 device tensor, x_t complex64 [B, F, T])."""
        assert X0.dtype == torch.complex64 and X0.shape == Y.shape and X0.dim() == 3 and X0.is_cuda
        B, F, T = X0.shape
        X0, Y = X0.contiguous(), Y.contiguous()
        t_host = t.detach().to("cpu", torch.float32).contiguous()
        coef = coef.detach().to("cpu", torch.float32).contiguous()
        assert t_host.numel() == B and tuple(coef.shape) == (2, B)
        gfp = self.net_module.gfp_features(t_host)
        x_t = torch.empty_like(X0)
        loss = torch.empty(1 + B, dtype=torch.float32, device=X0.device)
        nptr = None
        if noise is not None:
            noise = noise.to(torch.complex64).contiguous()
            assert noise.shape == X0.shape and noise.is_cuda
            nptr = noise.data_ptr()
        with torch.cuda.device(self.device):
            ws = self.workspace(B, F, T)
            _lib.check(self.L.use_train_forward(self.h, B, F, T, X0.data_ptr(), Y.data_ptr(), t_host.data_ptr(), gfp.data_ptr(),
                                                coef.data_ptr(), nptr, int(seed) & (2**64 - 1), int(clip0), int(bool(mae)),
                                                x_t.data_ptr(), loss.data_ptr(), ws.data_ptr(), ws.numel(),
                                                _lib.stream_ptr()), "use_train_forward")
        return loss, x_t

    def net_gfp(self, t_host):
        return self.net_module.gfp_features(t_host)

    def pc_sample(self, Y: torch.Tensor, ts: torch.Tensor, G: torch.Tensor, prior_std: float, noise=None, seed: int = 0,
                  clip0: int = 0, predictor="reverse_diffusion", corrector="none", corrector_steps=1, snr=0.5,
                  probability_flow=False, denoise=True, g=None, ald_step=None, trace=None, x_init=None, dt_steps=0, cond=None,
                  cond2=None):
        """The fused predictor-corrector loop (use_pc_sample_ex); returns (x_result, x_state), complex64 [B, F, T]:
        x_result = the noise-free mean of the last step (denoise) or the state."""
        assert Y.dtype == torch.complex64 and Y.dim() == 3 and Y.is_cuda
        B, F, T = Y.shape
        Y = Y.contiguous()
        N = int(ts.numel())
        ts = ts.detach().to("cpu", torch.float32).contiguous()
        G = G.detach().to("cpu", torch.float32).contiguous()
        gfp = self.net_module.gfp_features(ts)
        o = _lib.UseSamplerOpts()
        o.predictor, o.corrector = _lib.PRED[predictor], _lib.CORR[corrector]
        o.corrector_steps, o.snr = int(corrector_steps), float(snr)
        o.probability_flow, o.denoise, o.dt_steps = int(bool(probability_flow)), int(denoise), int(dt_steps)
        keep = []
        if g is not None:
            g = g.detach().to("cpu", torch.float32).contiguous()
            assert g.numel() == N
            o.g_host = g.data_ptr()
        if ald_step is not None:
            ald_step = ald_step.detach().to("cpu", torch.float32).contiguous()
            assert ald_step.numel() == N
            o.ald_step_host = ald_step.data_ptr()
        n_corr = 0 if corrector == "none" else int(corrector_steps)
        draws = 1 + N * (n_corr + (0 if predictor == "none" else 1))
        x_state, x_mean = torch.empty_like(Y), torch.empty_like(Y)
        nptr = None
        if noise is not None:
            assert noise.dtype == torch.complex64 and tuple(noise.shape) == (draws, B, F, T) and noise.is_cuda, \
                f"explicit noise must be complex64 [{draws}, {B}, {F}, {T}] (prior + per-step draws in the reference's order)"
            noise = noise.contiguous()
            nptr = noise.data_ptr()
        if trace is not None:
            assert trace.dtype == torch.complex64 and tuple(trace.shape) == (N, B, F, T) and trace.is_cuda \
                and trace.is_contiguous()
            o.trace = trace.data_ptr()
        if x_init is not None:
            x_init = x_init.to(torch.complex64).contiguous()
            assert tuple(x_init.shape) == (B, F, T) and x_init.is_cuda
            keep.append(x_init)
            o.x_init = x_init.data_ptr()
        if cond is not None:
            cond = cond.to(torch.complex64).contiguous()
            assert tuple(cond.shape) == (B, F, T) and cond.is_cuda
            keep.append(cond)
            o.cond = cond.data_ptr()
        if cond2 is not None:
            cond2 = cond2.to(torch.complex64).contiguous()
            assert tuple(cond2.shape) == (B, F, T) and cond2.is_cuda
            keep.append(cond2)
            o.cond2 = cond2.data_ptr()
        with torch.cuda.device(self.device):
            ws = self.workspace(B, F, T)
            _lib.check(self.L.use_pc_sample_ex(self.h, B, F, T, Y.data_ptr(), x_state.data_ptr(), x_mean.data_ptr(), N,
                                               ts.data_ptr(), G.data_ptr(), gfp.data_ptr(), float(prior_std), nptr,
                                               int(seed) & (2**64 - 1), int(clip0), C.byref(o), ws.data_ptr(), ws.numel(),
                                               _lib.stream_ptr()), "use_pc_sample_ex")
        return x_mean, x_state
