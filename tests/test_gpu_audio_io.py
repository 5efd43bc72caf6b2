"""Predict-side I/O on the GPU: FFT resampling with the semantics of the reference's
``librosa.resample(..., res_type="fft")`` (= scipy.signal.resample to ceil(n * ratio) samples, loadwav_dataset.py:94-98),
peak normalisation x0.8 (:99-100), zero padding to the longest clip (collate.py:42-73) and the asynchronous wav writer.
scipy.signal.resample in float64 is the checker (scipy is what librosa calls for res_type="fft"); the reference itself
computes in float64 and casts to float32, the GPU path computes in float32: tolerance 2e-5 of the clip's peak."""
import ctypes as C
import math
import os

import numpy as np
import pytest
import scipy.signal
import torch

from use_b200 import _lib
from use_b200.predict import GpuAudioPrep, LoadWavDataModule
from use_b200.sgmse_module import AsyncWavWriter
from oracle import sgmse_oracle as O
from util import stream

pytestmark = pytest.mark.gpu


def gpu_resample(x: np.ndarray, n_out: int) -> np.ndarray:
    L = _lib.lib()
    B, n_in = x.shape
    xd = torch.from_numpy(x).cuda()
    y = torch.zeros(B, n_out + 3, device="cuda")  # stride > n_out: the kernel must respect it
    need = C.c_size_t()
    assert L.use_resample_workspace_bytes(B, n_in, n_out, C.byref(need)) == 0
    work = torch.empty(need.value, dtype=torch.uint8, device="cuda")
    rc = L.use_resample_fft_f32(xd.data_ptr(), B, n_in, y.data_ptr(), n_out, n_out + 3, work.data_ptr(), work.numel(), stream())
    assert rc == 0, L.use_last_error()
    torch.cuda.synchronize()
    assert float(y[:, n_out:].abs().max()) == 0.0
    return y[:, :n_out].cpu().numpy()


@pytest.mark.parametrize("n_in,sr_in,sr_out", [
    (6400, 16000, 24000),    # up 3/2, even lengths
    (6401, 16000, 24000),    # odd input length
    (11025, 44100, 24000),   # down, irrational-looking ratio, m even/odd mix
    (9601, 48000, 24000),    # down 2x, odd (prime-ish) length -> Bluestein really needed
    (9600, 48000, 24000),    # down 2x, even: the unpaired middle bin is doubled
    (4800, 8000, 24000),     # up 3x
    (64000, 16000, 24000),   # a 4 s clip at 16 kHz -> the BASELINE clip length 96000
    (63997, 16000, 24000),   # prime input length
])
def test_fft_resample_matches_scipy(n_in, sr_in, sr_out):
    n_out = int(math.ceil(n_in * sr_out / sr_in))
    g = np.random.default_rng(n_in)
    t = np.arange(n_in) / sr_in
    x = np.stack([0.3 * g.standard_normal(n_in), 0.5 * np.sin(2 * np.pi * 440.0 * t) + 0.1 * g.standard_normal(n_in)]).astype(np.float32)
    ref = scipy.signal.resample(x.astype(np.float64), n_out, axis=-1)
    got = gpu_resample(x, n_out)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert got.shape == (2, n_out) and err < 2e-5, err


def test_peak_normalize_and_pad():
    L = _lib.lib()
    lens = [1000, 37, 640, 999]
    g = np.random.default_rng(0)
    y = g.standard_normal((4, 1000)).astype(np.float32) * np.array([[0.1], [2.0], [0.0], [1.0]], np.float32)
    yd = torch.from_numpy(y).cuda()
    ld = torch.tensor(lens, dtype=torch.int32, device="cuda")
    peaks = torch.empty(4, dtype=torch.int32, device="cuda")
    assert L.use_peak_normalize_pad_f32(yd.data_ptr(), ld.data_ptr(), 4, 1000, 0.8, peaks.data_ptr(), stream()) == 0
    torch.cuda.synchronize()
    got = yd.cpu().numpy()
    for b, n in enumerate(lens):
        peak = np.abs(y[b, :n]).max()
        ref = y[b, :n] / peak * 0.8 if peak > 0 else y[b, :n]
        assert np.allclose(got[b, :n], ref, rtol=2e-6, atol=1e-7) and np.all(got[b, n:] == 0.0)
        assert peaks.view(torch.float32)[b].item() == pytest.approx(float(peak), rel=1e-7)
    assert np.isfinite(got).all()  # the all-zero clip stays zero instead of 0/0


def _write_inputs(folder):
    from scipy.io import wavfile

    os.makedirs(os.path.join(folder, "sub"), exist_ok=True)
    files = {}
    a = (0.3 * O.synthetic_clips(1, 9600, seed=1)[0]).numpy()
    wavfile.write(os.path.join(folder, "mono24k.wav"), 24000, a)
    files["mono24k.wav"] = (a, 24000)
    st = np.stack([0.2 * O.synthetic_clips(1, 6401, seed=2)[0].numpy(), 0.9 * np.ones(6401, np.float32)], axis=1)
    wavfile.write(os.path.join(folder, "sub", "stereo16k.wav"), 16000, st)
    files[os.path.join("sub", "stereo16k.wav")] = (st[:, 0], 16000)
    i16 = (O.synthetic_clips(1, 11025, seed=3)[0].numpy() * 20000).astype(np.int16)
    wavfile.write(os.path.join(folder, "int16_44k.wav"), 44100, i16)
    files["int16_44k.wav"] = (i16.astype(np.float32) / 32768.0, 44100)
    return files


def test_datamodule_batches_match_the_reference_pipeline(tmp_path):
    """decode (host) -> channel 0 -> FFT resample -> x0.8 peak normalise -> pad to longest, batch born on the GPU, against
    the same pipeline restated with scipy / numpy in float64 (LoadWavDataset.__getitem__ + collate)."""
    src = str(tmp_path / "noisy")
    files = _write_inputs(src)
    dm = LoadWavDataModule(data_folder=src, target_folder=str(tmp_path / "out"), normalize=True, sampling_rate=24000,
                           batch_size=8)
    assert len(dm) == 3
    batches = list(dm.predict_dataloader())
    assert len(batches) == 1
    b = batches[0]
    assert b["perturbed"].is_cuda and b["perturbed"].dtype == torch.float32
    assert b["data_folder"] == src and b["sampling_rate"] == [24000] * 3
    got = b["perturbed"].cpu().numpy()
    for i, path in enumerate(b["audio_path"]):
        wav, sr = files[os.path.relpath(path, src)]
        ref = wav.astype(np.float64)
        if sr != 24000:
            ref = scipy.signal.resample(ref, int(math.ceil(len(ref) * 24000 / sr)))
        ref = ref / np.abs(ref).max() * 0.8
        n = int(b["sample_length"][i])
        assert n == len(ref) and b["name"][i] == os.path.basename(path)[:-4]
        assert np.abs(got[i, :n] - ref).max() < 5e-5 and np.all(got[i, n:] == 0.0)
    # batch_size 2 -> two batches, each padded to ITS longest clip
    dm2 = LoadWavDataModule(data_folder=src, target_folder=str(tmp_path / "out"), normalize=False, sampling_rate=None,
                            batch_size=2)
    bs = list(dm2.predict_dataloader())
    assert [x["perturbed"].shape[0] for x in bs] == [2, 1]
    assert all(int(x["sample_length"].max()) == x["perturbed"].shape[1] for x in bs)


def test_async_writer_round_trip(tmp_path):
    from scipy.io import wavfile

    w = AsyncWavWriter(workers=3)
    data = torch.randn(6, 5000, device="cuda") * 0.1
    for i in range(6):
        w.submit(str(tmp_path / "deep" / f"{i}.wav"), data[i, : 4000 + i], 24000)
    data.zero_()  # the writer must have captured the values in stream order before this
    w.close()
    ref = torch.randn(1)  # noqa: F841
    for i in range(6):
        sr, wav = wavfile.read(str(tmp_path / "deep" / f"{i}.wav"))
        assert sr == 24000 and wav.shape == (4000 + i,) and np.abs(wav).max() > 0
