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


class AsyncWavWriter:
    """Writes enhanced waveforms without stalling the sampler: ``submit`` enqueues a device->host copy into a pinned buffer
    on a side stream (ordered after the producing stream by an event) and hands the file write to a worker thread, which
    waits for the copy's event only.  ``close`` drains the queue.  The reference writes synchronously inside predict_step
    (SGMSE_module.py:71-80), i.e. the GPU idles during every ``sf.write``."""

    def __init__(self, workers: int = 4):
        from concurrent.futures import ThreadPoolExecutor

        self.pool = ThreadPoolExecutor(max(1, workers))
        self.futures = []
        self.copy_stream = None

    def submit(self, path: str, wav_dev: torch.Tensor, sample_rate: int) -> None:
        if not wav_dev.is_cuda:
            self.futures.append(self.pool.submit(self._write, path, wav_dev.detach().numpy().astype(np.float32), None, sample_rate))
            return
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=wav_dev.device)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(wav_dev.device))
        host = torch.empty(wav_dev.shape, dtype=torch.float32, pin_memory=True)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            host.copy_(wav_dev.detach(), non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        wav_dev.record_stream(self.copy_stream)
        self.futures.append(self.pool.submit(self._write, path, host, done, sample_rate))

    @staticmethod
    def _write(path, host, done, sample_rate):
        if done is not None:
            done.synchronize()
            host = host.numpy()
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        write_wav(path, host, int(sample_rate))
        return path

    def close(self):
        """Wait for every pending file; re-raises the first write error."""
        for f in self.futures:
            f.result()
        self.futures = []
        self.pool.shutdown(wait=True)


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
    def predict_step(self, batch: dict, batch_idx: int = 0, write: bool = True, writer: AsyncWavWriter = None) -> dict:
        """``writer``: an AsyncWavWriter -> files are written in the background (predict() joins it); without one the
        reference's synchronous behaviour is kept (the files exist when predict_step returns)."""
        batch = self.Score.sample(batch)
        if write and "audio_path" in batch:
            own = writer is None
            w = AsyncWavWriter(workers=2) if own else writer
            for i in range(len(batch["enhanced"])):
                noisy_path = batch["audio_path"][i]
                sample_length = int(batch["sample_length"][i])
                sample_rate = batch["sampling_rate"][i]
                enhanced_path = noisy_path.replace(batch["data_folder"], batch["target_folder"])
                w.submit(enhanced_path, batch["enhanced"][i, :sample_length], int(sample_rate))  # trim (:76-79)
            if own:
                w.close()
        return batch
