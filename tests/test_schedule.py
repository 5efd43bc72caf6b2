"""The integer step schedule: i in [0, N) -> (t_i, G_i) float32 bit patterns, bit-exact vs the reference."""
import os

import numpy as np
import torch

import use_b200
from oracle import sgmse_oracle as O
from util import GOLDEN


def test_step_tables_bit_exact():
    g = np.load(os.path.join(GOLDEN, "schedule.npz"))
    sde = use_b200.OUVESDE()
    for N in (3, 30, 50, 60):
        ts, G, std1 = sde.step_tables(N, 3e-2)
        assert ts.dtype == torch.float32 and G.dtype == torch.float32
        assert np.array_equal(ts.numpy().view(np.uint32), g[f"t_{N}"]), N
        assert np.array_equal(G.numpy().view(np.uint32), g[f"G_{N}"]), N
        assert np.float32(std1).view(np.uint32) == g["std1"][0]
        ots, oG = O.step_coefficients(N)
        assert torch.equal(ots, ts) and torch.equal(oG, G)


def test_schedule_is_monotone_and_hits_ends():
    ts, G, std1 = use_b200.OUVESDE().step_tables(30, 3e-2)
    assert float(ts[0]) == 1.0 and abs(float(ts[-1]) - 0.03) < 1e-7
    assert bool((ts[1:] < ts[:-1]).all()) and bool((G[1:] < G[:-1]).all())
    assert abs(std1 - 0.38898) < 1e-4


def test_ouve_matches_oracle_functions():
    sde = use_b200.OUVESDE()
    t = torch.tensor([1.0, 0.5, 0.03])
    assert torch.equal(sde._std(t), O.ouve_std(t))
    x = torch.zeros(3, 1, 2, 2, dtype=torch.complex64)
    _, g = sde.sde(x, t, x)
    assert torch.equal(g, O.ouve_diffusion(t))
