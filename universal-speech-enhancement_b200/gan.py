"""LSGAN refinement stage, predict side, on the B200 engine.

Mirrors /root/reference/src/models/components/GAN/generator/ncsnpp/model_wrapper.py:19-123 (``NCSNPP_Wrapper``:
STFT + compression -> ONE forward of ``NCSNpp(discriminative=True)`` -> decompression + iSTFT, result in
``batch["fake"]``) and the predict path of /root/reference/src/models/LSGAN_module.py:139-155 (``GANModule``).
Training (the "clean" branch, discriminators, criteria) is out of scope and raises.  It reuses every kernel of the
score path: the engine is configured with input_channels = 2, no time embedding, no 1/t scaling.
"""
from __future__ import annotations

import os
from math import ceil

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .backbones import NCSNpp
from .model_wrapper import get_window
from .sgmse_module import AsyncWavWriter, write_wav

try:  # pragma: no cover
    from lightning import LightningModule as _Base
except Exception:  # noqa: BLE001
    _Base = nn.Module


class NCSNPP_Wrapper(nn.Module):
    def __init__(self, n_fft=510, hop_length=128, num_frames=256, window="hann", spec_factor=0.15,
                 spec_abs_exponent=0.5, dtype: str = "fp32"):
        super().__init__()
        self.n_fft, self.hop_length, self.num_frames = n_fft, hop_length, num_frames
        self.window = get_window(window, n_fft)
        self.spec_factor, self.spec_abs_exponent = spec_factor, spec_abs_exponent
        self.target_len = (num_frames - 1) * hop_length
        self.dtype_name = dtype
        self.net = NCSNpp(discriminative=True, compute_dtype=dtype)
        self.net._spec = dict(n_fft=n_fft, hop_length=hop_length, spec_factor=spec_factor,
                              spec_abs_exponent=spec_abs_exponent)
        self._tables = {}

    def _dev_tables(self, device, Tp):
        key = (str(device), Tp)
        tb = self._tables.get(key)
        if tb is None:
            n = self.n_fft
            ang = 2.0 * np.pi * torch.arange(n, dtype=torch.float64) / n
            tw = torch.stack([torch.cos(ang), torch.sin(ang)], dim=1).to(torch.float32)
            w64 = self.window.to(torch.float64)
            env = torch.zeros(n + self.hop_length * (Tp - 1), dtype=torch.float64)
            for f in range(Tp):
                env[f * self.hop_length: f * self.hop_length + n] += w64 * w64
            tb = dict(window=self.window.to(device=device, dtype=torch.float32).contiguous(),
                      twiddle=tw.contiguous().to(device), env=env.to(torch.float32).to(device))
            self._tables = {key: tb}
        return tb

    @torch.no_grad()
    def forward(self, batch_data: dict) -> dict:
        if "clean" in batch_data:
            raise NotImplementedError("the training branch of NCSNPP_Wrapper is out of scope of the B200 predict path")
        y = batch_data["perturbed"]
        if not y.is_cuda:
            raise RuntimeError("NCSNPP_Wrapper(B200) runs on CUDA tensors only; there is no CPU path")
        y = y.to(torch.float32).contiguous()
        B, L = y.shape
        T = 1 + L // self.hop_length
        Tp = int(ceil(T / 64) * 64)
        F = self.n_fft // 2 + 1
        eng = self.net.engine(y.device, self.dtype_name)
        tb = self._dev_tables(y.device, Tp)
        Y = torch.empty(B, F, Tp, dtype=torch.complex64, device=y.device)
        with torch.cuda.device(y.device):
            _lib.check(eng.L.use_stft(eng.h, B, L, Tp, y.data_ptr(), Y.data_ptr(), tb["window"].data_ptr(),
                                      tb["twiddle"].data_ptr(), _lib.stream_ptr()), "use_stft")
            X = eng.net(Y, None, None)
            out = torch.empty(B, L, dtype=torch.float32, device=y.device)
            frames = torch.empty(B, Tp, self.n_fft, dtype=torch.float32, device=y.device)
            _lib.check(eng.L.use_istft(eng.h, B, L, Tp, X.data_ptr(), out.data_ptr(), frames.data_ptr(),
                                       tb["window"].data_ptr(), tb["twiddle"].data_ptr(), tb["env"].data_ptr(),
                                       _lib.stream_ptr()), "use_istft")
        batch_data["fake"] = out
        return batch_data


class GANModule(_Base):
    """Predict-side ``GANModule``: ``predict_step`` runs the generator and writes ``batch["fake"]`` (LSGAN_module.py:139-155).
    The lenient ``load_state_dict`` of the reference (skips missing / mis-shaped tensors, :51-61) is kept."""

    def __init__(self, G: nn.Module, D: nn.Module = None, G_optimizer=None, D_optimizer=None, G_scheduler=None,
                 D_scheduler=None, G_criterion=None, D_criterion=None, compile: bool = False,
                 accumulate_grad_batches: int = 1, rewrite_lr=False, G_lr=None, D_lr=None) -> None:
        super().__init__()
        self.G = G
        self.D = D
        self.compile = compile

    def load_state_dict(self, state_dict, strict=True):
        own = self.state_dict()
        for name, param in state_dict.items():
            if name in own and own[name].size() == param.size():
                own[name].copy_(param)
        for m in self.modules():
            if hasattr(m, "invalidate_engine"):
                m.invalidate_engine()

    def load_checkpoint(self, ckpt_path: str, strict: bool = False):
        """Load a Lightning ``.ckpt`` of the reference (``state_dict`` keys ``G.net.*`` / ``D.*``) through the lenient
        ``load_state_dict`` above -- what ``trainer.predict(ckpt_path=...)`` does for this module (predict.py:79)."""
        ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
        return self.load_state_dict(ckpt.get("state_dict", ckpt), strict=strict)

    def training_step(self, batch, batch_idx):
        raise NotImplementedError("training is out of scope of the B200 predict path")

    @torch.no_grad()
    def predict_step(self, batch: dict, batch_idx: int = 0, write: bool = True, writer=None) -> dict:
        batch = self.G(batch)
        if write and "audio_path" in batch:
            own = writer is None
            w = AsyncWavWriter(workers=2) if own else writer
            for i in range(len(batch["fake"])):
                noisy_path = batch["audio_path"][i]
                n = int(batch["sample_length"][i])
                out_path = noisy_path.replace(batch["data_folder"], batch["target_folder"])
                w.submit(out_path, batch["fake"][i, :n], int(batch["sampling_rate"][i]))
            if own:
                w.close()
        return batch
