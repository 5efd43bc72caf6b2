"""CPU oracle for the SGMSE reverse-SDE sampling path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``use_b200``) never routes through it and fails loudly when the CUDA library is missing.

It restates, in plain functional PyTorch (dtype generic: float32 for parity, float64 for tolerance
budgeting), the algorithm the reference runs for ``src/predict.py``:

    ScoreModel.sample            /root/reference/src/models/components/sgmse/model_wrapper.py:262-329
    sampling.get_pc_sampler      /root/reference/src/models/components/sgmse/sampling/__init__.py:59-71
    ReverseDiffusionPredictor    /root/reference/src/models/components/sgmse/sampling/predictors.py:61-68
    SDE.discretize / RSDE        /root/reference/src/models/components/sgmse/sdes.py:75-92,159-173
    OUVESDE.sde/_std/prior       /root/reference/src/models/components/sgmse/sdes.py:216-254
    NCSNpp.forward               /root/reference/src/models/components/sgmse/backbones/ncsnpp.py:324-501
    layerspp blocks              /root/reference/src/models/components/sgmse/backbones/ncsnpp_utils/layerspp.py
    FIR resampling               /root/reference/src/models/components/sgmse/backbones/ncsnpp_utils/up_or_down_sampling.py:188-264
                                 /root/reference/src/models/components/sgmse/backbones/ncsnpp_utils/op/upfirdn2d.py:173-208

Parity pin: the reference's own tests hold no golden vector for this path, so
the pin is the reference implementation itself, imported unmodified in the build container by
``oracle/make_golden.py``; that script asserts this restatement reproduces ``ScoreModel.sample``
bit-for-bit in float32 and writes ``tests/golden/*.npz``, which ``tests/test_oracle_golden.py``
re-checks everywhere (the GPU box has no /root/reference).

Third-party arithmetic (conv2d, group_norm, stft, softmax) lives in PyTorch for the reference too
(requirements.txt pins torch==2.3.0; goldens were generated under the torch of this image).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class NetCfg:
    """Architecture hyper-parameters of NCSNpp (ncsnpp.py:45-68); defaults = NCSNppLarge (:511-518)."""

    nf: int = 128
    ch_mult: Tuple[int, ...] = (1, 1, 2, 2, 2, 2, 2)
    num_res_blocks: int = 2
    input_channels: int = 4
    fourier_scale: float = 16.0
    fir_kernel: Tuple[int, ...] = (1, 3, 3, 1)
    # discriminative=True (the GAN generator, ncsnpp.py:88-94) switches both off and uses input_channels=2
    conditional: bool = True
    scale_by_sigma: bool = True

    @property
    def num_resolutions(self) -> int:
        return len(self.ch_mult)


@dataclass(frozen=True)
class SpecCfg:
    """STFT / spectral-compression parameters (configs/model/SGMSE_Large.yaml:11-16)."""

    n_fft: int = 1022
    hop_length: int = 160
    spec_factor: float = 0.15
    spec_abs_exponent: float = 0.5
    window: str = "hann"


@dataclass(frozen=True)
class SdeCfg:
    """OUVESDE constants (sdes.py:184) and t_eps (SGMSE_Large.yaml:7)."""

    theta: float = 1.5
    sigma_min: float = 0.05
    sigma_max: float = 0.5
    t_eps: float = 3e-2
    T: float = 1.0

    @property
    def logsig(self) -> float:
        return float(np.log(self.sigma_max / self.sigma_min))


LARGE = NetCfg()
TINY = NetCfg(nf=64, ch_mult=(1, 2), num_res_blocks=1)
# NCSNpp(discriminative=True) with the class defaults: the LSGAN generator (GAN/generator/ncsnpp/model_wrapper.py:54)
LARGE6 = NetCfg(nf=128, ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2, input_channels=6)  # condition="both"
GAN_G = NetCfg(nf=128, ch_mult=(1, 2, 2, 2), num_res_blocks=1, input_channels=2, conditional=False, scale_by_sigma=False)
GAN_TINY = NetCfg(nf=64, ch_mult=(1, 2), num_res_blocks=1, input_channels=2, conditional=False, scale_by_sigma=False)


# --------------------------------------------------------------------------------------------
# module list structure (ncsnpp.py:186-316) and seeded non-degenerate weights
# --------------------------------------------------------------------------------------------
def module_plan(cfg: NetCfg) -> List[dict]:
    """The ``all_modules`` list in construction order: kind + channel info per entry."""
    nf, nres = cfg.nf, cfg.num_resolutions
    plan: List[dict] = [{"kind": "gfp"}]
    if cfg.conditional:
        plan += [{"kind": "linear", "cin": 2 * nf, "cout": 4 * nf}, {"kind": "linear", "cin": 4 * nf, "cout": 4 * nf}]
    plan.append({"kind": "conv3", "cin": cfg.input_channels, "cout": nf})
    hs_c = [nf]
    in_ch = nf
    for lvl in range(nres):
        for _ in range(cfg.num_res_blocks):
            out_ch = nf * cfg.ch_mult[lvl]
            plan.append({"kind": "rb", "cin": in_ch, "cout": out_ch, "up": False, "down": False})
            in_ch = out_ch
            hs_c.append(in_ch)
        if lvl != nres - 1:
            plan.append({"kind": "rb", "cin": in_ch, "cout": in_ch, "up": False, "down": True})
            plan.append({"kind": "combine", "cin": cfg.input_channels, "cout": in_ch})
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    plan.append({"kind": "rb", "cin": in_ch, "cout": in_ch, "up": False, "down": False})
    plan.append({"kind": "attn", "c": in_ch})
    plan.append({"kind": "rb", "cin": in_ch, "cout": in_ch, "up": False, "down": False})
    for lvl in reversed(range(nres)):
        for _ in range(cfg.num_res_blocks + 1):
            out_ch = nf * cfg.ch_mult[lvl]
            plan.append({"kind": "rb", "cin": in_ch + hs_c.pop(), "cout": out_ch, "up": False, "down": False})
            in_ch = out_ch
        plan.append({"kind": "gn", "c": in_ch})
        plan.append({"kind": "conv3", "cin": in_ch, "cout": cfg.input_channels})
        if lvl != 0:
            plan.append({"kind": "rb", "cin": in_ch, "cout": in_ch, "up": True, "down": False})
    assert not hs_c
    return plan


def make_state_dict(cfg: NetCfg = LARGE, seed: int = 7, dtype=torch.float32) -> Dict[str, Tensor]:
    """Seeded, NON-degenerate weights with the reference's names and shapes.

    The reference default init scales every Conv_1 / NIN_3 / pyramid conv by 1e-10
    (layers.py:100-103, ncsnpp.py:59) and zeroes all biases, which would leave half the kernels
    numerically unexercised; no checkpoint is available offline.  Here every
    weight is N(0, 1/fan_in), every GroupNorm weight 1 + 0.1 N(0,1), every bias 0.1 N(0,1).
    Keys are relative to ``score_net`` ("all_modules.{i}...." / "output_layer....").
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def randn(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    def conv(prefix, cout, cin, k):
        sd[prefix + ".weight"] = randn(cout, cin, k, k) / math.sqrt(cin * k * k)
        sd[prefix + ".bias"] = 0.1 * randn(cout)

    def lin(prefix, cout, cin):
        sd[prefix + ".weight"] = randn(cout, cin) / math.sqrt(cin)
        sd[prefix + ".bias"] = 0.1 * randn(cout)

    def gn(prefix, c):
        sd[prefix + ".weight"] = 1.0 + 0.1 * randn(c)
        sd[prefix + ".bias"] = 0.1 * randn(c)

    def nin(prefix, c):
        sd[prefix + ".W"] = randn(c, c) / math.sqrt(c)
        sd[prefix + ".b"] = 0.1 * randn(c)

    conv("output_layer", 2, cfg.input_channels, 1)
    for i, m in enumerate(module_plan(cfg)):
        p = f"all_modules.{i}"
        k = m["kind"]
        if k == "gfp":
            sd[p + ".W"] = randn(cfg.nf) * cfg.fourier_scale
        elif k == "linear":
            lin(p, m["cout"], m["cin"])
        elif k == "conv3":
            conv(p, m["cout"], m["cin"], 3)
        elif k == "gn":
            gn(p, m["c"])
        elif k == "combine":
            conv(p + ".Conv_0", m["cout"], m["cin"], 1)
        elif k == "attn":
            gn(p + ".GroupNorm_0", m["c"])
            for j in range(4):
                nin(p + f".NIN_{j}", m["c"])
        elif k == "rb":
            gn(p + ".GroupNorm_0", m["cin"])
            conv(p + ".Conv_0", m["cout"], m["cin"], 3)
            lin(p + ".Dense_0", m["cout"], 4 * cfg.nf)
            gn(p + ".GroupNorm_1", m["cout"])
            conv(p + ".Conv_1", m["cout"], m["cout"], 3)
            if m["cin"] != m["cout"] or m["up"] or m["down"]:
                conv(p + ".Conv_2", m["cout"], m["cin"], 1)
    return {k: v.to(dtype) for k, v in sd.items()}


# --------------------------------------------------------------------------------------------
# FIR resampling (up_or_down_sampling.py:188-264 -> op/upfirdn2d.py:173-208, the CPU branch)
# --------------------------------------------------------------------------------------------
def _setup_kernel(k: Sequence[float]) -> np.ndarray:
    k = np.asarray(k, dtype=np.float32)
    if k.ndim == 1:
        k = np.outer(k, k)
    k /= np.sum(k)
    return k


def _upfirdn2d(x: Tensor, kernel: Tensor, up: int, down: int, pad0: int, pad1: int) -> Tensor:
    """zero-insert upsample -> pad -> correlate with the flipped kernel -> decimate."""
    _, channel, in_h, in_w = x.shape
    x = x.reshape(-1, in_h, in_w, 1)
    kh, kw = kernel.shape
    out = x.view(-1, in_h, 1, in_w, 1, 1)
    out = F.pad(out, [0, 0, 0, up - 1, 0, 0, 0, up - 1])
    out = out.view(-1, in_h * up, in_w * up, 1)
    out = F.pad(out, [0, 0, max(pad0, 0), max(pad1, 0), max(pad0, 0), max(pad1, 0)])
    out = out.permute(0, 3, 1, 2)
    out = out.reshape([-1, 1, in_h * up + pad0 + pad1, in_w * up + pad0 + pad1])
    w = torch.flip(kernel, [0, 1]).view(1, 1, kh, kw)
    out = F.conv2d(out, w)
    out = out.reshape(-1, 1, in_h * up + pad0 + pad1 - kh + 1, in_w * up + pad0 + pad1 - kw + 1)
    out = out.permute(0, 2, 3, 1)
    out = out[:, ::down, ::down, :]
    out_h = (in_h * up + pad0 + pad1 - kh) // down + 1
    out_w = (in_w * up + pad0 + pad1 - kw) // down + 1
    return out.view(-1, channel, out_h, out_w)


def fir_upsample_2d(x: Tensor, k: Sequence[float] = (1, 3, 3, 1), factor: int = 2) -> Tensor:
    """upsample_2d (up_or_down_sampling.py:202-232).  The reference always builds a float32 tap
    tensor; for the float64 budget runs the taps follow x.dtype (they are exact in either)."""
    kk = _setup_kernel(k) * (factor**2)
    p = kk.shape[0] - factor
    return _upfirdn2d(x, torch.tensor(kk, dtype=x.dtype), factor, 1, (p + 1) // 2 + factor - 1, p // 2)


def fir_downsample_2d(x: Tensor, k: Sequence[float] = (1, 3, 3, 1), factor: int = 2) -> Tensor:
    """downsample_2d (up_or_down_sampling.py:235-264)."""
    kk = _setup_kernel(k)
    p = kk.shape[0] - factor
    return _upfirdn2d(x, torch.tensor(kk, dtype=x.dtype), 1, factor, (p + 1) // 2, p // 2)


# --------------------------------------------------------------------------------------------
# NCSN++ forward (ncsnpp.py:324-501, layerspp.py)
# --------------------------------------------------------------------------------------------
def _gn(sd, p: str, x: Tensor) -> Tensor:
    c = x.shape[1]
    return F.group_norm(x, min(c // 4, 32), sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _conv(sd, p: str, x: Tensor, pad: int) -> Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def _nin(sd, p: str, x: Tensor) -> Tensor:
    # layers.py:639-650
    y = torch.einsum("bhwc,cd->bhwd", x.permute(0, 2, 3, 1), sd[p + ".W"]) + sd[p + ".b"]
    return y.permute(0, 3, 1, 2)


def resblock(sd, p: str, m: dict, x: Tensor, temb: Tensor, fir_k) -> Tensor:
    """ResnetBlockBigGANpp.forward (layerspp.py:282-314)."""
    h = F.silu(_gn(sd, p + ".GroupNorm_0", x))
    if m["up"]:
        h = fir_upsample_2d(h, fir_k)
        x = fir_upsample_2d(x, fir_k)
    elif m["down"]:
        h = fir_downsample_2d(h, fir_k)
        x = fir_downsample_2d(x, fir_k)
    h = _conv(sd, p + ".Conv_0", h, 1)
    if temb is not None:
        h = h + F.linear(F.silu(temb), sd[p + ".Dense_0.weight"], sd[p + ".Dense_0.bias"])[:, :, None, None]
    h = F.silu(_gn(sd, p + ".GroupNorm_1", h))
    h = _conv(sd, p + ".Conv_1", h, 1)
    if m["cin"] != m["cout"] or m["up"] or m["down"]:
        x = _conv(sd, p + ".Conv_2", x, 0)
    return (x + h) / np.sqrt(2.0)


def attnblock(sd, p: str, x: Tensor) -> Tensor:
    """AttnBlockpp.forward (layerspp.py:77-93)."""
    B, C, H, W = x.shape
    h = _gn(sd, p + ".GroupNorm_0", x)
    q = _nin(sd, p + ".NIN_0", h)
    k = _nin(sd, p + ".NIN_1", h)
    v = _nin(sd, p + ".NIN_2", h)
    w = torch.einsum("bchw,bcij->bhwij", q, k) * (int(C) ** (-0.5))
    w = torch.reshape(w, (B, H, W, H * W))
    w = F.softmax(w, dim=-1)
    w = torch.reshape(w, (B, H, W, H, W))
    h = torch.einsum("bhwij,bcij->bchw", w, v)
    h = _nin(sd, p + ".NIN_3", h)
    return (x + h) / np.sqrt(2.0)


def time_embedding(sd, t: Tensor) -> Tensor:
    """GaussianFourierProjection(log t) -> Linear -> SiLU -> Linear (layerspp.py:30-39, ncsnpp.py:349-368)."""
    x = torch.log(t)
    x_proj = x[:, None] * sd["all_modules.0.W"][None, :] * 2 * np.pi
    temb = torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)
    temb = F.linear(temb, sd["all_modules.1.weight"], sd["all_modules.1.bias"])
    temb = F.linear(F.silu(temb), sd["all_modules.2.weight"], sd["all_modules.2.bias"])
    return temb


def ncsnpp_forward(sd: Dict[str, Tensor], cfg: NetCfg, x: Tensor, t: Optional[Tensor], taps: Optional[dict] = None) -> Tensor:
    """x: complex [B, 2, F, T] (= cat[x_t, Y]);  t: [B].  Returns complex [B, 1, F, T].
    Discriminative configs (cfg.conditional False): x complex [B, 1, F, T], t None.

    ``taps`` (optional dict) receives intermediate real tensors for per-layer debugging of the CUDA path.
    """
    plan = module_plan(cfg)
    fir_k = cfg.fir_kernel
    # complex -> [Re x, Im x, Re Y, Im Y]  (ncsnpp.py:333-347)
    xr = torch.cat([torch.cat([x[:, [c]].real, x[:, [c]].imag], dim=1) for c in range(cfg.input_channels // 2)], dim=1)
    temb = time_embedding(sd, t) if cfg.conditional else None
    xr = 2 * xr - 1.0
    input_pyramid = xr
    i = 3 if cfg.conditional else 1
    hs = [_conv(sd, f"all_modules.{i}", xr, 1)]
    i += 1
    nres = cfg.num_resolutions
    for lvl in range(nres):
        for _ in range(cfg.num_res_blocks):
            h = resblock(sd, f"all_modules.{i}", plan[i], hs[-1], temb, fir_k)
            i += 1
            hs.append(h)
        if lvl != nres - 1:
            h = resblock(sd, f"all_modules.{i}", plan[i], hs[-1], temb, fir_k)
            i += 1
            input_pyramid = fir_downsample_2d(input_pyramid, fir_k)
            h = _conv(sd, f"all_modules.{i}.Conv_0", input_pyramid, 0) + h  # Combine "sum" (layerspp.py:50-55)
            i += 1
            hs.append(h)
    if taps is not None:
        taps["down_out"] = hs[-1]
    h = hs[-1]
    h = resblock(sd, f"all_modules.{i}", plan[i], h, temb, fir_k)
    i += 1
    h = attnblock(sd, f"all_modules.{i}", h)
    i += 1
    h = resblock(sd, f"all_modules.{i}", plan[i], h, temb, fir_k)
    i += 1
    if taps is not None:
        taps["mid_out"] = h
    pyramid = None
    for lvl in reversed(range(nres)):
        for _ in range(cfg.num_res_blocks + 1):
            h = resblock(sd, f"all_modules.{i}", plan[i], torch.cat([h, hs.pop()], dim=1), temb, fir_k)
            i += 1
        ph = F.silu(_gn(sd, f"all_modules.{i}", h))
        i += 1
        ph = _conv(sd, f"all_modules.{i}", ph, 1)
        i += 1
        if lvl == nres - 1:
            pyramid = ph
        else:
            pyramid = fir_upsample_2d(pyramid, fir_k)
            pyramid = pyramid + ph
        if lvl != 0:
            h = resblock(sd, f"all_modules.{i}", plan[i], h, temb, fir_k)
            i += 1
    assert not hs and i == len(plan)
    if taps is not None:
        taps["pyramid"] = pyramid
    h = pyramid
    if cfg.scale_by_sigma:
        h = pyramid / t.reshape(-1, 1, 1, 1)  # scale_by_sigma divides by the TIME value (ncsnpp.py:492-494)
    h = _conv(sd, "output_layer", h, 0)
    h = torch.reshape(h, (h.size(0), 2, 1, h.size(2), h.size(3)))
    h = torch.permute(h, (0, 2, 3, 4, 1)).contiguous()
    return torch.view_as_complex(h)


# --------------------------------------------------------------------------------------------
# STFT front / back end (model_wrapper.py:14-20,92-122; util/other.py:128-135)
# --------------------------------------------------------------------------------------------
def get_window(spec: SpecCfg, dtype=torch.float32) -> Tensor:
    w = torch.hann_window(spec.n_fft, periodic=True, dtype=dtype)
    if spec.window == "sqrthann":
        w = torch.sqrt(w)
    elif spec.window != "hann":
        raise NotImplementedError(spec.window)
    return w


def stft(sig: Tensor, spec: SpecCfg) -> Tensor:
    return torch.stft(sig, n_fft=spec.n_fft, hop_length=spec.hop_length, window=get_window(spec, sig.dtype),
                      center=True, return_complex=True)


def istft(S: Tensor, spec: SpecCfg, length: Optional[int]) -> Tensor:
    return torch.istft(S, n_fft=spec.n_fft, hop_length=spec.hop_length, window=get_window(spec, S.real.dtype),
                       center=True, length=length)


def spec_fwd(S: Tensor, spec: SpecCfg) -> Tensor:
    if spec.spec_abs_exponent != 1:
        e = spec.spec_abs_exponent
        S = S.abs() ** e * torch.exp(1j * S.angle())
    return S * spec.spec_factor


def spec_back(S: Tensor, spec: SpecCfg) -> Tensor:
    S = S / spec.spec_factor
    if spec.spec_abs_exponent != 1:
        e = spec.spec_abs_exponent
        S = S.abs() ** (1 / e) * torch.exp(1j * S.angle())
    return S


def pad_spec(Y: Tensor) -> Tensor:
    T = Y.size(3)
    num_pad = (64 - T % 64) if T % 64 != 0 else 0
    return F.pad(Y, (0, num_pad, 0, 0))


# --------------------------------------------------------------------------------------------
# SDE + sampler (sdes.py, sampling/__init__.py, predictors.py)
# --------------------------------------------------------------------------------------------
def timesteps(N: int, sde: SdeCfg = SdeCfg()) -> Tensor:
    """The float32 step schedule; must be BIT-exact (sampling/__init__.py:63)."""
    return torch.linspace(sde.T, sde.t_eps, N)


def ouve_std(t: Tensor, sde: SdeCfg = SdeCfg()) -> Tensor:
    """OUVESDE._std (sdes.py:231-243)."""
    sigma_min, theta, logsig = sde.sigma_min, sde.theta, sde.logsig
    return torch.sqrt(
        (sigma_min**2 * torch.exp(-2 * theta * t) * (torch.exp(2 * (theta + logsig) * t) - 1) * logsig)
        / (theta + logsig)
    )


def ouve_diffusion(t: Tensor, sde: SdeCfg = SdeCfg()) -> Tensor:
    """g(t) of OUVESDE.sde (sdes.py:216-224)."""
    sigma = sde.sigma_min * (sde.sigma_max / sde.sigma_min) ** t
    return sigma * np.sqrt(2 * sde.logsig)


def step_coefficients(N: int, sde: SdeCfg = SdeCfg()) -> Tuple[Tensor, Tensor]:
    """(t_i, G_i) float32 tables: G_i = g(t_i) * sqrt(float32(1/N))  (sdes.py:88-92, dt = 1/N)."""
    ts = timesteps(N, sde)
    G = ouve_diffusion(ts, sde) * torch.sqrt(torch.tensor(1 / N))
    return ts, G


def draw_noise(shape, N: int, seed: int, dtype=torch.complex64) -> Tensor:
    """The N+1 complex normal draws the reference makes (prior + one per step, the last unused by
    x_mean), in its order, from a seeded CPU generator: [N+1, *shape]."""
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randn(shape, dtype=dtype, generator=g) for _ in range(N + 1)])


def draws_per_step(predictor: str, corrector: str, corrector_steps: int) -> int:
    """Normal draws one outer step consumes: the corrector's inner steps, then the predictor's."""
    return (0 if corrector == "none" else corrector_steps) + (0 if predictor == "none" else 1)


def pc_sample_spec(score_fn, Y: Tensor, N: int, noise: Tensor, sde: SdeCfg = SdeCfg(),
                   trace: Optional[list] = None, predictor: str = "reverse_diffusion", corrector: str = "none",
                   corrector_steps: int = 1, snr: float = 0.5, denoise: bool = True) -> Tensor:
    """pc_sampler() (sampling/__init__.py:59-71) with explicit noise.

    score_fn(x, t_vec) -> complex score [B,1,F,T] (already negated net output).
    Y complex [B,1,F,T]; noise complex [1 + N * draws_per_step, B,1,F,T] in the order the reference draws:
    prior, then per outer step the corrector's inner-step draws followed by the predictor's draw.
    predictor: reverse_diffusion (predictors.py:56-68) | euler_maruyama (predictors.py:40-53 with the drift of
    RSDE.rsde_parts, sdes.py:128-150) | none.  corrector: none | langevin (correctors.py:37-64) | ald (:67-98).
    Returns x_mean of the last step (denoise) or the state.
    """
    B = Y.shape[0]
    rdt = Y.real.dtype
    bc = lambda v: v[:, None, None, None]  # noqa: E731
    std1 = ouve_std(torch.ones((B,), dtype=rdt), sde)
    xt = Y + noise[0] * bc(std1)
    ts = torch.linspace(sde.T, sde.t_eps, N, dtype=rdt)
    xt_mean = xt
    k = 1  # next unused draw
    n_corr = 0 if corrector == "none" else corrector_steps
    for i in range(N):
        vec_t = torch.ones(B, dtype=rdt) * ts[i]
        # ---- corrector
        if n_corr:
            std_t = ouve_std(vec_t, sde)
        for _ in range(n_corr):
            grad = score_fn(xt, vec_t)
            z = noise[k]
            k += 1
            if corrector == "langevin":
                gnorm = torch.norm(grad.reshape(B, -1), dim=-1).mean()
                znorm = torch.norm(z.reshape(B, -1), dim=-1).mean()
                step = ((snr * znorm / gnorm) ** 2 * 2).reshape(1, 1, 1, 1)
            elif corrector == "ald":
                step = bc((snr * std_t) ** 2 * 2)
            else:
                raise NotImplementedError(corrector)
            xt_mean = xt + step * grad
            xt = xt_mean + z * torch.sqrt(step * 2)
        # ---- predictor
        dt = 1 / N
        g = ouve_diffusion(vec_t, sde)
        if predictor == "reverse_diffusion":
            # SDE.discretize (sdes.py:88-92) with dt = 1/N
            drift = sde.theta * (Y - xt)
            G = g * torch.sqrt(torch.tensor(dt, dtype=rdt))
            f = drift * dt
            Gb = bc(G)
            rev_f = f - Gb**2 * score_fn(xt, vec_t) * 1.0
            z = noise[k]
            k += 1
            xt_mean = xt - rev_f
            xt = xt_mean + Gb * z
        elif predictor == "euler_maruyama":
            z = noise[k]
            k += 1
            gb = bc(g)
            total = sde.theta * (Y - xt) + (-(gb**2) * score_fn(xt, vec_t) * 1.0)
            xt_mean = xt + total * (-dt)
            xt = xt_mean + gb * np.sqrt(dt) * z
        elif predictor == "none":
            xt_mean = xt
        else:
            raise NotImplementedError(predictor)
        if trace is not None:
            trace.append(xt_mean.clone())
    return xt_mean if (denoise and N) else xt


def ode_sample_spec(score_fn, Y: Tensor, N: int, noise0: Tensor, sde: SdeCfg = SdeCfg(), rtol: float = 1e-5,
                    atol: float = 1e-5, method: str = "RK45", denoise: bool = True):
    """get_ode_sampler() (sampling/__init__.py:76-159): the probability-flow ODE dx/dt = theta (Y - x) - g(t)^2 score / 2
    (RSDE.sde with probability_flow=True, sdes.py:122-150) integrated from T to t_eps by scipy's solve_ivp over the
    flattened complex64 state, then one noise-free ReverseDiffusionPredictor step at t = t_eps (dt = 1/N).
    noise0: the prior draw, complex [B,1,F,T].  Returns (x, number of function evaluations)."""
    from scipy import integrate

    B = Y.shape[0]
    rdt = Y.real.dtype
    std1 = ouve_std(torch.ones((B,), dtype=rdt), sde)
    x = Y + noise0 * std1[:, None, None, None]

    def ode_func(t, x_flat):
        xt = torch.from_numpy(x_flat.reshape(tuple(Y.shape))).type(torch.complex64)
        vec_t = torch.ones(B) * t
        g = ouve_diffusion(vec_t, sde)[:, None, None, None]
        drift = sde.theta * (Y - xt) + (-(g**2) * score_fn(xt, vec_t) * 0.5)
        return drift.detach().numpy().reshape((-1,))

    sol = integrate.solve_ivp(ode_func, (sde.T, sde.t_eps), x.detach().numpy().reshape((-1,)), rtol=rtol, atol=atol,
                              method=method)
    x = torch.tensor(sol.y[:, -1]).reshape(Y.shape).type(torch.complex64)
    if denoise:
        vec_eps = torch.ones(B) * sde.t_eps
        dt = 1 / N
        G = (ouve_diffusion(vec_eps, sde) * torch.sqrt(torch.tensor(dt, dtype=rdt)))[:, None, None, None]
        f = sde.theta * (Y - x) * dt
        x = x - (f - G**2 * score_fn(x, vec_eps) * 1.0)
    return x, sol.nfev


def sample(sd: Dict[str, Tensor], y: Tensor, N: int, noise: Optional[Tensor] = None, seed: int = 42,
           net: NetCfg = LARGE, spec: SpecCfg = SpecCfg(), sde: SdeCfg = SdeCfg(),
           return_spec: bool = False, fake: Optional[Tensor] = None, condition: str = "noisy", sde_input: str = "noisy",
           **sampler_kw):
    """ScoreModel.sample (model_wrapper.py:262-329): y float [B, L] -> enhanced float [B, L].  ``fake`` (the GAN stage's
    output, batch["fake"]) + condition / sde_input in {"noisy", "denoised"} select the network's conditioning spectrogram
    and the SDE's y (:281-299)."""
    with torch.no_grad():
        T_orig = y.size(1)
        Y = pad_spec(spec_fwd(stft(y, spec), spec).unsqueeze(1))
        Yd = pad_spec(spec_fwd(stft(fake, spec), spec).unsqueeze(1)) if fake is not None else None
        cond = [Yd] if (condition == "denoised" and Yd is not None) else ([Y, Yd] if condition == "both" else [Y])
        Y = Yd if (sde_input == "denoised" and Yd is not None) else Y
        if noise is None:
            per = draws_per_step(sampler_kw.get("predictor", "reverse_diffusion"), sampler_kw.get("corrector", "none"),
                                 sampler_kw.get("corrector_steps", 1))
            noise = draw_noise(tuple(Y.shape), N * per, seed, dtype=Y.dtype)

        def score_fn(x, t):
            return -ncsnpp_forward(sd, net, torch.cat([x] + cond, dim=1), t)

        if sampler_kw.get("sampler_type") == "ode":
            xm, _ = ode_sample_spec(score_fn, Y, N, noise[0], sde, rtol=sampler_kw.get("rtol", 1e-5),
                                    atol=sampler_kw.get("atol", 1e-5))
        else:
            sampler_kw.pop("sampler_type", None)
            xm = pc_sample_spec(score_fn, Y, N, noise, sde, **sampler_kw)
        out = istft(spec_back(xm.squeeze(1), spec), spec, T_orig)
    return (out, xm, Y) if return_spec else out


def ouve_mean(x0: Tensor, t: Tensor, y: Tensor, sde: SdeCfg = SdeCfg()) -> Tensor:
    """OUVESDE._mean (sdes.py:225-228)."""
    e = torch.exp(-sde.theta * t)[:, None, None, None]
    return e * x0 + (1 - e) * y


def train_draws(B: int, n_freq: int, n_frames: int, seed: int, crop_range: Optional[int] = None, sde: SdeCfg = SdeCfg()):
    """The three random draws of ScoreModel.train_step under np.random.seed(seed) / torch.manual_seed(seed), in the
    reference's order AND memory layout: crop offset int(np.random.uniform(0, crop_range)) (model_wrapper.py:156; 0 when
    the clip is padded instead), t = rand(B) (T - t_eps) + t_eps (:178), z = randn_like(x) (:180).  randn_like fills its
    result in MEMORY order and x = spec_fwd(stft(.)) keeps torch.stft's frame-major strides (F*T, 1, F), so z is drawn
    frame by frame, not bin by bin.  Returns (start, t [B], z complex64 [B,1,F,T])."""
    start = 0
    if crop_range is not None:
        np.random.seed(seed)
        start = int(np.random.uniform(0, crop_range))
    g = torch.Generator().manual_seed(seed)
    t = torch.rand(B, generator=g) * (sde.T - sde.t_eps) + sde.t_eps
    z = torch.empty_strided((B, 1, n_freq, n_frames), (n_freq * n_frames, n_freq, 1, n_freq), dtype=torch.complex64)
    z.normal_(generator=g)
    return start, t, z


def train_step_loss(sd: Dict[str, Tensor], x: Tensor, y: Tensor, t: Tensor, z: Tensor, start: int = 0,
                    net: NetCfg = LARGE, spec: SpecCfg = SpecCfg(), sde: SdeCfg = SdeCfg(), num_frames: int = 512,
                    loss_type: str = "mse"):
    """Forward half of ScoreModel.train_step (model_wrapper.py:147-208) with the random draws made explicit:
    ``start`` (crop offset, np.random.uniform :156), ``t`` [B] (torch.rand :178), ``z`` complex [B,1,F,T] (:180).
    x (clean), y (noisy): float [B, L].  Returns (loss, perturbed x_t)."""
    with torch.no_grad():
        target_len = (num_frames - 1) * spec.hop_length
        cur = x.size(-1)
        pad = max(target_len - cur, 0)
        if pad == 0:
            x, y = x[..., start:start + target_len], y[..., start:start + target_len]
        else:
            x = F.pad(x, (pad // 2, pad // 2 + (pad % 2)))
            y = F.pad(y, (pad // 2, pad // 2 + (pad % 2)))
        X0 = spec_fwd(stft(x, spec), spec).unsqueeze(1)
        Y = spec_fwd(stft(y, spec), spec).unsqueeze(1)
        std = ouve_std(t, sde)[:, None, None, None]
        x_t = ouve_mean(X0, t, Y, sde) + std * z
        score = -ncsnpp_forward(sd, net, torch.cat([x_t, Y], dim=1), t)
        err = score * std + z
        per = err.abs() if loss_type == "mae" else torch.square(err.abs())
        loss = torch.mean(0.5 * torch.sum(per.reshape(per.shape[0], -1), dim=-1))
    return loss, x_t


def gan_denoise(sd: Dict[str, Tensor], y: Tensor, net: NetCfg = GAN_G, spec: SpecCfg = SpecCfg()) -> Tensor:
    """NCSNPP_Wrapper.forward, inference branch (GAN/generator/ncsnpp/model_wrapper.py:114-121): one forward of the
    discriminative NCSN++ on the compressed spectrogram: y float [B, L] -> batch["fake"] float [B, L]."""
    with torch.no_grad():
        T_orig = y.size(1)
        Y = pad_spec(spec_fwd(stft(y, spec), spec).unsqueeze(1))
        out = ncsnpp_forward(sd, net, Y, None)
        return istft(spec_back(out.squeeze(1), spec), spec, T_orig)


def synthetic_clips(B: int, L: int = 96000, seed: int = 1234) -> Tensor:
    """The bench / parity input: clamp(0.1 randn, -1, 1)."""
    g = torch.Generator().manual_seed(seed)
    return torch.clamp(0.1 * torch.randn(B, L, generator=g), -1.0, 1.0)
