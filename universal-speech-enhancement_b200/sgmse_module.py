"""SGMSEModule: the predict-side LightningModule of the reference, on the B200 path.

Drop-in for /root/reference/src/models/SGMSE_module.py:10-108 as far as ``src/predict.py`` uses it:
constructor ``(Score, optimizer, scheduler, compile)``, checkpoint keys ``Score.score_net.*`` and
``predict_step(batch, batch_idx)`` = ``Score.sample(batch)`` + device->host + trim to ``sample_length`` + wav
write (:65-82).  Training hooks are out of scope and raise.
``lightning`` is not installed in the build image, so the base class degrades to ``torch.nn.Module``.
"""
from __future__ import annotations

import os

import numpy as np
import torch

try:  # pragma: no cover - lightning is absent offline
    from lightning import LightningModule as _Base
except Exception:  # noqa: BLE001
    _Base = torch.nn.Module


def write_wav(path: str, wav: np.ndarray, sample_rate: int) -> None:
    try:
        import soundfile as sf  # the reference's writer (SGMSE_module.py:80)

        sf.write(path, wav, sample_rate)
    except ImportError:
        from scipy.io import wavfile

        wavfile.write(path, int(sample_rate), wav.astype(np.float32))


class SGMSEModule(_Base):
    def __init__(self, Score: torch.nn.Module, optimizer=None, scheduler=None, compile: bool = False) -> None:
        super().__init__()
        self.Score = Score
        self.optimizer = optimizer
        self.scheduler = scheduler
        self.compile = compile

    def configure_optimizers(self):
        raise NotImplementedError("training is out of scope of the B200 sampling path")

    def training_step(self, batch, batch_idx):
        raise NotImplementedError("training is out of scope of the B200 sampling path")

    validation_step = test_step = training_step

    def load_checkpoint(self, ckpt_path: str, strict: bool = True):
        """Load a Lightning ``.ckpt`` of the reference (``state_dict`` keys ``Score.score_net.*``)."""
        ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
        sd = ckpt.get("state_dict", ckpt)
        return self.load_state_dict(sd, strict=strict)

    @torch.no_grad()
    def predict_step(self, batch: dict, batch_idx: int = 0, write: bool = True) -> dict:
        batch = self.Score.sample(batch)
        if write and "audio_path" in batch:
            for i in range(len(batch["enhanced"])):
                noisy_path = batch["audio_path"][i]
                sample_length = int(batch["sample_length"][i])
                sample_rate = batch["sampling_rate"][i]
                enhanced_path = noisy_path.replace(batch["data_folder"], batch["target_folder"])
                os.makedirs(os.path.dirname(enhanced_path) or ".", exist_ok=True)
                wav = batch["enhanced"][i].detach().cpu().numpy().astype(np.float32)[:sample_length]
                write_wav(enhanced_path, wav, int(sample_rate))
        return batch
