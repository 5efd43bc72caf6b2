"""Predict entrypoint: the caller side of the path, kept minimal.

Mirrors /root/reference/src/predict.py:39-92 (`predict(cfg)`: instantiate datamodule + model, load the checkpoint,
run predict over the dataloader) and the batch contract of the reference's LoadWavDataModule + collate
(src/data/components/loadwav_dataset.py:90-120, collate.py:42-73): dict with `perturbed` f32 [B, Lmax] zero padded,
`sample_length`, `sampling_rate`, `audio_path`, `name`, `data_folder`, `target_folder`.

    python -m use_b200.predict model=SGMSE_Large ckpt_path=... data.data_folder=... data.target_folder=...
"""
from __future__ import annotations

import os
import sys
from typing import Dict, Iterator, List

import numpy as np
import torch

from .config import compose, instantiate


def _read_wav(path: str):
    try:
        import soundfile as sf

        wav, sr = sf.read(path, dtype="float32")
    except ImportError:
        from scipy.io import wavfile

        sr, wav = wavfile.read(path)
        if wav.dtype.kind == "i":
            wav = wav.astype(np.float32) / float(np.iinfo(wav.dtype).max)
        wav = wav.astype(np.float32)
    if wav.ndim > 1:
        wav = wav[:, 0]  # the reference keeps channel 0 (loadwav_dataset.py:96), it does not down-mix
    return np.ascontiguousarray(wav), int(sr)


class LoadWavDataModule:
    """Walk `data_folder` for .wav files, mono, resample to `sampling_rate`, peak-normalise x0.8, pad to the longest."""

    def __init__(self, data_folder: str, target_folder: str, normalize: bool = True, sampling_rate: int = 24000,
                 batch_size: int = 1, num_workers: int = 0):
        self.data_folder, self.target_folder = data_folder, target_folder
        self.normalize, self.sampling_rate, self.batch_size = normalize, sampling_rate, batch_size
        self.files: List[str] = []
        for root, _, names in os.walk(data_folder):
            self.files += [os.path.join(root, n) for n in sorted(names) if n.lower().endswith(".wav")]

    def _load(self, path: str) -> np.ndarray:
        wav, sr = _read_wav(path)
        if sr != self.sampling_rate:
            from scipy.signal import resample

            wav = resample(wav, int(round(len(wav) * self.sampling_rate / sr))).astype(np.float32)
        if self.normalize:
            peak = float(np.abs(wav).max())
            wav = wav / peak * 0.8 if peak > 0 else wav  # loadwav_dataset.py:99-100; an all-zero file stays zero
        return wav

    def predict_dataloader(self) -> Iterator[Dict]:
        for i in range(0, len(self.files), self.batch_size):
            paths = self.files[i:i + self.batch_size]
            wavs = [self._load(p) for p in paths]
            lens = [len(w) for w in wavs]
            batch = np.zeros((len(wavs), max(lens)), dtype=np.float32)
            for j, w in enumerate(wavs):
                batch[j, : len(w)] = w
            yield {"perturbed": torch.from_numpy(batch), "sample_length": torch.tensor(lens, dtype=torch.int32),
                   "sampling_rate": [self.sampling_rate] * len(wavs), "audio_path": paths,
                   "name": [os.path.basename(p) for p in paths], "data_folder": self.data_folder,
                   "target_folder": self.target_folder}


def predict(cfg: Dict):
    if not cfg.get("ckpt_path") and not cfg.get("allow_random_init"):
        raise AssertionError("ckpt_path is required (predict.py:48 of the reference asserts the same)")
    datamodule = instantiate(cfg["data"])
    model = instantiate(cfg["model"])
    if cfg.get("ckpt_path"):
        model.load_checkpoint(cfg["ckpt_path"])
    dev = torch.device("cuda", torch.cuda.current_device())
    outs = []
    for i, batch in enumerate(datamodule.predict_dataloader()):
        batch["perturbed"] = batch["perturbed"].to(dev, non_blocking=True)
        outs.append(model.predict_step(batch, i))
    return outs


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = compose(os.path.join(root, "configs"), "predict.yaml", argv)
    return predict(cfg)


if __name__ == "__main__":
    main()
