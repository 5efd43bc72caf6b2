"""Golden for the forward half of ScoreModel.train_step, from the UNMODIFIED reference.

TEST INFRASTRUCTURE (build container only: needs /root/reference).

Runs the reference's ``ScoreModel.train_step({"clean": x, "perturbed": y})`` (model_wrapper.py:147-208) under
``np.random.seed(s)`` / ``torch.manual_seed(s)`` with the oracle's seeded NCSNppLarge weights (B = 2 clips of 90 000
samples, so the random crop to (512 - 1) * 160 samples is exercised), then replays the same three random draws
explicitly (crop offset, t ~ U(t_eps, 1), z) through oracle.train_step_loss and asserts BIT-exact equality of the loss.
Writes tests/golden/train_step_T512.npz (loss for "mse" and "mae", the draws, a sub-sampled x_t).

    python oracle/make_golden_train.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import sgmse_oracle as O  # noqa: E402
from oracle.make_golden import import_reference  # noqa: E402


def clips(B, L):
    clean = O.synthetic_clips(B, L, seed=21)
    return clean, clean + 0.5 * O.synthetic_clips(B, L, seed=22)


def main():
    torch.set_num_threads(os.cpu_count())
    ScoreModel, _, _, _ = import_reference()
    sdL = O.make_state_dict(O.LARGE, seed=7)
    B, L, seed = 2, 90000, 5
    x, y = clips(B, L)
    out = dict(B=B, L=L, seed=seed, weight_seed=7)
    for loss_type in ("mse", "mae"):
        m = ScoreModel(backbone="ncsnpplarge", sde="ouve", t_eps=3e-2, mode="regen-joint-training", condition="noisy",
                       loss_type=loss_type, n_fft=1022, hop_length=160, num_frames=512, window="hann", spec_factor=0.15,
                       spec_abs_exponent=0.5, sde_input="noisy").eval()
        m.score_net.load_state_dict(sdL, strict=True)
        np.random.seed(seed)
        torch.manual_seed(seed)
        with torch.no_grad():
            ref = m.train_step({"clean": x.clone(), "perturbed": y.clone()})
        # replay the draws
        start, t, z = O.train_draws(B, 512, 512, seed, crop_range=L - m.target_len)
        mine, x_t = O.train_step_loss(sdL, x, y, t, z, start=start, loss_type=loss_type)
        print(f"{loss_type}: reference loss {float(ref):.9g}  oracle {float(mine):.9g}  start {start}  t {t.tolist()}", flush=True)
        assert torch.equal(ref, mine), loss_type
        out[f"loss_{loss_type}"] = np.float32(ref.item())
        out["start"], out["t"] = start, t.numpy()
        out["xt_re"], out["xt_im"] = x_t.real[:, 0, ::16, ::16].numpy(), x_t.imag[:, 0, ::16, ::16].numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "train_step_T512.npz"), **out)
    print("written tests/golden/train_step_T512.npz")


if __name__ == "__main__":
    main()
