import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, use_b200
from oracle import sgmse_oracle as O
m = use_b200.ScoreModel(backbone="ncsnpplarge", sde="ouve", t_eps=3e-2, condition="noisy", sde_input="noisy", n_fft=1022, hop_length=160, num_frames=512, dtype="bf16")
m.score_net.load_state_dict(O.make_state_dict(O.LARGE, seed=7), strict=True)
y = O.synthetic_clips(3, 96000).cuda()
Y = m.stft_compressed(y).unsqueeze(1)
g = torch.Generator().manual_seed(1)
x = Y + 0.3 * torch.randn(Y.shape, dtype=torch.complex64, generator=g).cuda()
t = torch.tensor([0.5, 0.5, 0.5]).cuda()
full = m(x, t, [Y], Y)
res = []
for b in range(3):
    one = m(x[b:b+1], t[b:b+1], [Y[b:b+1]], Y[b:b+1])
    res.append(bool(torch.equal(one, full[b:b+1])))
print(os.environ.get("USE_B200_CONV_NSPLIT_MAXTILES"), "score alone==batch:", res)
