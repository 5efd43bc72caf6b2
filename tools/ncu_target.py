"""Short, profiler-friendly invocation of the hot path: one STFT, `--evals` network evaluations + SDE steps of
NCSNppLarge on `--batch` full-size clips (4 s @ 24 kHz -> 512 x 640), one iSTFT.  Used under `ncu` (never timed)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import use_b200  # noqa: E402
from oracle import sgmse_oracle as O  # noqa: E402  (seeded weights / clips generator only)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--evals", type=int, default=1)
ap.add_argument("--dtype", default="bf16")
a = ap.parse_args()
m = use_b200.ScoreModel(backbone="ncsnpplarge", sde="ouve", t_eps=3e-2, condition="noisy", sde_input="noisy", n_fft=1022,
                        hop_length=160, num_frames=512, dtype=a.dtype)
m.score_net.load_state_dict(O.make_state_dict(O.LARGE, seed=7), strict=True)
y = O.synthetic_clips(a.batch, 96000).cuda()
out = m.sample({"perturbed": y}, N=a.evals, seed=1)["enhanced"]
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.abs().mean()))
