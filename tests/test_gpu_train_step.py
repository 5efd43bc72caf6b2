"""Forward half of ScoreModel.train_step (model_wrapper.py:147-208) on the B200 path
(use_train_forward: marginal_prob perturbation -> one score evaluation with PER-SAMPLE times -> denoising-score-matching
loss) against the golden of the UNMODIFIED reference (tests/golden/train_step_T512.npz, oracle/make_golden_train.py).
Tolerance on the scalar loss: 1e-3 relative (fp32 / TF32), 1e-2 (bf16); x_t (no network involved) 1e-6."""
import os

import numpy as np
import pytest
import torch

import use_b200
from oracle import sgmse_oracle as O
from util import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu


def _clips(B, L):
    clean = O.synthetic_clips(B, L, seed=21)
    return clean, clean + 0.5 * O.synthetic_clips(B, L, seed=22)


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-3), ("bf16", 1e-2)])
@pytest.mark.parametrize("loss_type", ["mse", "mae"])
def test_train_step_forward_matches_reference_golden(loss_type, dtype, tol):
    g = np.load(os.path.join(GOLDEN, "train_step_T512.npz"))
    B, L, seed = int(g["B"]), int(g["L"]), int(g["seed"])
    m = use_b200.ScoreModel(backbone="ncsnpplarge", sde="ouve", t_eps=3e-2, condition="noisy", sde_input="noisy",
                            loss_type=loss_type, n_fft=1022, hop_length=160, num_frames=512, dtype=dtype)
    m.score_net.load_state_dict(O.make_state_dict(O.LARGE, seed=int(g["weight_seed"])), strict=True)
    x, y = _clips(B, L)
    start, t, z = O.train_draws(B, 512, 512, seed, crop_range=L - 511 * 160)  # the reference's draws, its memory layout
    assert np.array_equal(t.numpy(), g["t"]) and start == int(g["start"])
    np.random.seed(seed)  # train_step draws the crop offset itself, like the reference (np.random.uniform)
    loss, parts = m.train_step({"clean": x.cuda(), "perturbed": y.cuda()}, t=t, noise=z.cuda(), return_parts=True)
    ref = float(g[f"loss_{loss_type}"])
    assert loss.shape == () and abs(float(loss) - ref) <= tol * abs(ref), (float(loss), ref)
    xt = parts["x_t"][:, 0, ::16, ::16].cpu()
    ref_xt = torch.complex(torch.from_numpy(g["xt_re"]), torch.from_numpy(g["xt_im"]))
    assert rel_l2(torch.view_as_real(xt), torch.view_as_real(ref_xt)) < 1e-5
    assert abs(float(parts["per_clip"].mean()) - float(loss)) <= 1e-6 * abs(ref)


def test_train_step_random_draws_and_padding():
    """Without explicit draws: t from torch's RNG, z from Philox(seed) -- deterministic given the seeds; short clips are
    centre-padded to (num_frames - 1) * hop samples (model_wrapper.py:160-165)."""
    m = use_b200.ScoreModel(backbone="ncsnpplarge", sde="ouve", t_eps=3e-2, condition="noisy", sde_input="noisy",
                            n_fft=1022, hop_length=160, num_frames=64, dtype="bf16")
    m.score_net.load_state_dict(O.make_state_dict(O.LARGE, seed=7), strict=True)
    x, y = _clips(3, 8000)  # < 63 * 160 = 10080 -> padded
    torch.manual_seed(1)
    a = m.train_step({"clean": x.cuda(), "perturbed": y.cuda()}, seed=4)
    torch.manual_seed(1)
    b = m.train_step({"clean": x.cuda(), "perturbed": y.cuda()}, seed=4)
    assert torch.equal(a, b) and bool(torch.isfinite(a)) and float(a) > 0
    torch.manual_seed(1)
    c = m.train_step({"clean": x.cuda(), "perturbed": y.cuda()}, seed=5)
    assert not torch.equal(a, c)
