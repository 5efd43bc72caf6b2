"""Pin parity on the BASELINE configuration itself: one 4 s @ 24 kHz clip (512 x 640 spectrogram), N = 30.

TEST INFRASTRUCTURE (build container only: needs /root/reference, which does not travel to the GPU box).

Runs the UNMODIFIED reference ``ScoreModel.sample({"perturbed": y}, N=30)`` (model_wrapper.py:262-329) with the
oracle's seeded weights, observing every ``ReverseDiffusionPredictor.update_fn`` return value (predictors.py:61-68)
through a call-time wrapper (the reference file is not modified), then runs the oracle on the same input and asserts
max |reference - oracle| == 0.0 for the waveform and for x_mean at every step.  Writes

    tests/golden/sample_large_T640_N30.npz
        y            float32 [1, 96000]     the clip (== oracle.synthetic_clips(1, 96000, seed=1234))
        enhanced     float32 [1, 96000]     reference output
        xmean_re/im  float32 [30, 64, 128]  x_mean after every step, sub-sampled [::8, ::5] of the 512 x 640 grid
        xnorm        float64 [30]           ||x_mean||_2 of the FULL grid per step (to turn errors into rel-L2)

≈ 6-7 min per run on 8 cores (two runs: reference, oracle).

    python oracle/make_golden_large.py
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import sgmse_oracle as O  # noqa: E402
from oracle.make_golden import import_reference  # noqa: E402

SUB_F, SUB_T = 8, 5


def main():
    torch.set_num_threads(os.cpu_count())
    ScoreModel, _, _, _ = import_reference()
    from src.models.components.sgmse.sampling import predictors as P  # type: ignore

    sdL = O.make_state_dict(O.LARGE, seed=7)
    m = ScoreModel(backbone="ncsnpplarge", sde="ouve", t_eps=3e-2, mode="regen-joint-training", condition="noisy",
                   loss_type="mse", n_fft=1022, hop_length=160, num_frames=512, window="hann", spec_factor=0.15,
                   spec_abs_exponent=0.5, sde_input="noisy").eval()
    m.score_net.load_state_dict(sdL, strict=True)
    B, L, N, seed = 1, 96000, 30, 42
    y = O.synthetic_clips(B, L)

    ref_trace = []
    orig = P.ReverseDiffusionPredictor.update_fn

    def observed(self, x, t, *a, **kw):
        out = orig(self, x, t, *a, **kw)
        ref_trace.append(out[1].clone())
        return out

    P.ReverseDiffusionPredictor.update_fn = observed
    t0 = time.time()
    torch.manual_seed(seed)
    ref = m.sample({"perturbed": y.clone()}, N=N)["enhanced"]
    P.ReverseDiffusionPredictor.update_fn = orig
    print(f"reference sample(): {time.time() - t0:.1f} s, {len(ref_trace)} steps", flush=True)
    assert len(ref_trace) == N

    t0 = time.time()
    spec = O.SpecCfg()
    with torch.no_grad():
        Y = O.pad_spec(O.spec_fwd(O.stft(y, spec), spec).unsqueeze(1))
        noise = O.draw_noise(tuple(Y.shape), N, seed, dtype=Y.dtype)
        trace = []
        xm = O.pc_sample_spec(lambda x, t: -O.ncsnpp_forward(sdL, O.LARGE, torch.cat([x, Y], 1), t), Y, N, noise,
                              trace=trace)
        mine = O.istft(O.spec_back(xm.squeeze(1), spec), spec, L)
    print(f"oracle sample(): {time.time() - t0:.1f} s", flush=True)
    d = float((ref - mine).abs().max())
    ds = max(float((a - b).abs().max()) for a, b in zip(ref_trace, trace))
    print("max|ref-oracle| waveform =", d, " x_mean over all steps =", ds, flush=True)
    assert d == 0.0 and ds == 0.0

    xs = torch.stack([x[0, 0] for x in ref_trace])  # [N, 512, 640] complex
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "sample_large_T640_N30.npz"),
        y=y.numpy(), enhanced=ref.numpy(),
        xmean_re=xs.real[:, ::SUB_F, ::SUB_T].contiguous().numpy(),
        xmean_im=xs.imag[:, ::SUB_F, ::SUB_T].contiguous().numpy(),
        xnorm=np.array([float(torch.view_as_real(x).double().norm()) for x in xs]),
        xnorm_sub=np.array([float(torch.view_as_real(x[::SUB_F, ::SUB_T].contiguous()).double().norm()) for x in xs]),
        B=B, L=L, N=N, seed=seed, weight_seed=7, sub_f=SUB_F, sub_t=SUB_T)
    print("written tests/golden/sample_large_T640_N30.npz")


if __name__ == "__main__":
    main()
