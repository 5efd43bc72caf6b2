import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, use_b200
from oracle import sgmse_oracle as O
m = use_b200.ScoreModel(backbone="ncsnpplarge", sde="ouve", t_eps=3e-2, condition="noisy", sde_input="noisy", n_fft=1022, hop_length=160, num_frames=512, dtype="bf16")
m.score_net.load_state_dict(O.make_state_dict(O.LARGE, seed=7), strict=True)
y4 = O.synthetic_clips(4, 96000).cuda()
def run(sl, clip0, N=2): return m.sample({"perturbed": y4[sl]}, N=N, seed=5, clip0=clip0)["enhanced"]
a4 = run(slice(0,4), 0)
a4b = run(slice(0,4), 0)
print("B4 deterministic", torch.equal(a4, a4b))
for i in range(4):
    bi = run(slice(i,i+1), i)
    bi2 = run(slice(i,i+1), i)
    print("clip", i, "alone==alone", torch.equal(bi, bi2), "alone==B4", torch.equal(bi, a4[i:i+1]))
p01 = run(slice(0,2), 0); p23 = run(slice(2,4), 2)
print("pair01==B4", torch.equal(p01, a4[:2]), "pair23==B4", torch.equal(p23, a4[2:]))
eng = m.score_net.engine(y4.device, "bf16")
eng.set_option("overlap_groups", 1)
a4g1 = run(slice(0,4), 0)
print("B4 one group == B4 two groups", torch.equal(a4g1, a4), " onegroup==pairs", torch.equal(a4g1[:2], p01), torch.equal(a4g1[2:], p23))
