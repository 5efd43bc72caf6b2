"""Predict entrypoint and the I/O either side of the sampler.

Mirrors /root/reference/src/predict.py:39-92 (``predict(cfg)``: instantiate datamodule + model, load the checkpoint,
run predict over the dataloader) and the batch contract of the reference's LoadWavDataModule + collate
(src/data/loadwav_datamodule.py:12-74, src/data/components/loadwav_dataset.py:11-120, collate.py:42-73): dict with
``perturbed`` f32 [B, Lmax] zero padded, ``sample_length`` int32 [B], ``sampling_rate``, ``audio_path``, ``name``,
``data_folder``, ``target_folder``.

What is different from the reference's CPU workers: once the sampler runs at ~10 clips/s per GPU the single-threaded
host loop (decode -> FFT resample -> normalise -> pad -> ... -> write) is the wall-clock bottleneck of ``predict``, so

  * the files of a batch are DECODED on the host (a thread pool; decoding is I/O) and staged in pinned memory,
  * resampling (librosa ``res_type="fft"`` = scipy.signal.resample semantics, any length), channel-0 selection having
    happened at decode time, peak normalisation to 0.8 and padding to the longest clip run ON THE GPU
    (``use_resample_fft_f32`` / ``use_peak_normalize_pad_f32``, csrc/resample.cu); the batch is born on the device,
  * the next batch is prepared (decode + H2D) by a background thread while the current one is being sampled,
  * enhanced waveforms leave through ``AsyncWavWriter`` (sgmse_module.py): D2H into pinned buffers on a side stream,
    files written by worker threads; ``predict`` joins the writer at the end.

    python -m use_b200.predict model=SGMSE_Large ckpt_path=... data.data_folder=... data.target_folder=...
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import queue
import sys
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Iterator, List, Optional

import numpy as np
import torch

from . import _lib
from .config import compose, instantiate


def read_wav(path: str):
    """-> (float32 mono [n] = channel 0, sample rate).  The reference keeps channel 0 (loadwav_dataset.py:95-96)."""
    try:
        import soundfile as sf

        wav, sr = sf.read(path, dtype="float32")
    except ImportError:
        from scipy.io import wavfile

        sr, wav = wavfile.read(path)
        if wav.dtype.kind == "i":
            wav = wav.astype(np.float32) / float(np.iinfo(wav.dtype).max + 1)
        elif wav.dtype.kind == "u":  # 8-bit PCM is unsigned
            wav = (wav.astype(np.float32) - 128.0) / 128.0
        wav = wav.astype(np.float32)
    if wav.ndim > 1:
        wav = wav[:, 0]
    return np.ascontiguousarray(wav), int(sr)


class GpuAudioPrep:
    """resample -> peak-normalise -> pad of one batch of decoded clips, on the GPU (C ABI, no torch math)."""

    def __init__(self, device: torch.device, sampling_rate: Optional[int], normalize: bool):
        self.device, self.sampling_rate, self.normalize = device, sampling_rate, normalize
        self.L = _lib.lib()
        self._work = None

    def out_length(self, n: int, sr: int) -> int:
        if not self.sampling_rate or sr == self.sampling_rate:
            return n
        return int(math.ceil(n * float(self.sampling_rate) / sr))  # librosa.resample: ceil(len * ratio)

    def __call__(self, wavs: List[np.ndarray], srs: List[int]) -> (torch.Tensor, List[int]):
        lens = [self.out_length(len(w), sr) for w, sr in zip(wavs, srs)]
        B, Lmax = len(wavs), max(lens)
        with torch.cuda.device(self.device):
            batch = torch.empty(B, Lmax, dtype=torch.float32, device=self.device)
            st = _lib.stream_ptr()
            for b, (w, sr, n_out) in enumerate(zip(wavs, srs, lens)):
                src = torch.from_numpy(w).pin_memory().to(self.device, non_blocking=True)
                if n_out == len(w):
                    batch[b, :n_out].copy_(src)
                    continue
                need = C.c_size_t()
                _lib.check(self.L.use_resample_workspace_bytes(1, len(w), n_out, C.byref(need)), "use_resample_workspace_bytes")
                if self._work is None or self._work.numel() < need.value:
                    self._work = torch.empty(need.value, dtype=torch.uint8, device=self.device)
                _lib.check(self.L.use_resample_fft_f32(src.data_ptr(), 1, len(w), batch[b].data_ptr(), n_out, Lmax,
                                                       self._work.data_ptr(), self._work.numel(), st), "use_resample_fft_f32")
            ld = torch.tensor(lens, dtype=torch.int32).to(self.device, non_blocking=True)
            peaks = torch.empty(B, dtype=torch.int32, device=self.device)
            _lib.check(self.L.use_peak_normalize_pad_f32(batch.data_ptr(), ld.data_ptr(), B, Lmax, 0.8 if self.normalize else 0.0,
                                                         peaks.data_ptr(), st), "use_peak_normalize_pad_f32")
        return batch, lens


class LoadWavDataModule:
    """Same constructor keys as the reference's LoadWavDataModule (loadwav_datamodule.py:13-30); file discovery as
    LoadWavDataset.__init__ (loadwav_dataset.py:38-77: json lines / list file / in-memory lists / folder walk)."""

    def __init__(self, list_path=None, json_path=None, data_folder=None, input_json_list=None, input_plain_list=None,
                 normalize: bool = False, min_duration_seconds=None, max_duration_seconds=None, sampling_rate=None,
                 output_resample=False, output_resample_rate=None, target_folder=None, batch_size: int = 64,
                 num_workers: int = 0, pin_memory: bool = False, prefetch: int = 1):
        if output_resample:
            raise NotImplementedError("output_resample (a second resampling of the network INPUT, loadwav_dataset.py:101-104) "
                                      "is not on the accelerated path")
        self.data_folder, self.target_folder = data_folder, target_folder
        self.normalize, self.sampling_rate, self.batch_size = normalize, sampling_rate, batch_size
        self.num_workers, self.prefetch = max(1, int(num_workers) or 4), max(0, int(prefetch))
        self.files: List[str] = []

        def add_json(line):
            line = line.strip()
            if not line:
                return
            d = json.loads(line)
            p = d.get("audio_filepath", d.get("file_path"))
            if p not in self.files:
                self.files.append(p)

        if json_path:
            with open(json_path) as f:
                for line in f:
                    add_json(line)
        elif list_path:
            with open(list_path) as f:
                self.files += [line.strip() for line in f if line.strip()]
        elif input_json_list:
            for line in input_json_list:
                add_json(line)
        elif input_plain_list:
            self.files += list(input_plain_list)
        elif data_folder:
            for root, _, names in os.walk(data_folder):
                self.files += [os.path.join(root, n) for n in names if n.endswith(".wav")]
        else:
            raise ValueError("No input list provided")

    def __len__(self):
        return len(self.files)

    def _decode(self, paths: List[str], pool: ThreadPoolExecutor):
        out = list(pool.map(read_wav, paths))
        return [w for w, _ in out], [sr for _, sr in out]

    def predict_dataloader(self, device: Optional[torch.device] = None) -> Iterator[Dict]:
        """Yields batches whose ``perturbed`` already lives on ``device`` (default: the current CUDA device)."""
        device = device or torch.device("cuda", torch.cuda.current_device())
        prep = GpuAudioPrep(device, self.sampling_rate, self.normalize)
        chunks = [self.files[i:i + self.batch_size] for i in range(0, len(self.files), self.batch_size)]
        pool = ThreadPoolExecutor(self.num_workers)
        q: "queue.Queue" = queue.Queue(maxsize=max(1, self.prefetch))

        def producer():  # decode ahead of the GPU: the sampler of batch i overlaps the file I/O of batch i + 1
            try:
                for paths in chunks:
                    q.put((paths, self._decode(paths, pool)))
            except BaseException as e:  # noqa: BLE001 - surfaced in the consumer
                q.put(e)
            q.put(None)

        threading.Thread(target=producer, daemon=True).start()
        while True:
            item = q.get()
            if item is None:
                break
            if isinstance(item, BaseException):
                raise item
            paths, (wavs, srs) = item
            batch, lens = prep(wavs, srs)
            rates = [self.sampling_rate if self.sampling_rate else sr for sr in srs]
            out = {"perturbed": batch, "sample_length": torch.tensor(lens, dtype=torch.int32), "sampling_rate": rates,
                   "audio_path": paths, "name": [os.path.basename(p).split(".wav")[0] for p in paths]}
            if self.data_folder:
                out["data_folder"] = self.data_folder
            if self.target_folder:
                out["target_folder"] = self.target_folder
            yield out
        pool.shutdown(wait=False)


def predict(cfg: Dict):
    if not cfg.get("ckpt_path") and not cfg.get("allow_random_init"):
        raise AssertionError("ckpt_path is required (predict.py:48 of the reference asserts the same)")
    datamodule = instantiate(cfg["data"])
    model = instantiate(cfg["model"])
    if cfg.get("ckpt_path"):
        model.load_checkpoint(cfg["ckpt_path"])
    dev = torch.device("cuda", torch.cuda.current_device())
    from .sgmse_module import AsyncWavWriter

    writer = AsyncWavWriter(workers=4)
    outs = []
    try:
        for i, batch in enumerate(datamodule.predict_dataloader(dev)):
            outs.append(model.predict_step(batch, i, writer=writer))
    finally:
        writer.close()
    return outs


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = compose(os.path.join(root, "configs"), "predict.yaml", argv)
    return predict(cfg)


if __name__ == "__main__":
    main()
