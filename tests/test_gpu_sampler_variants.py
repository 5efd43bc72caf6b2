"""The other registered sampler variants on the fused CUDA loop (use_pc_sample_ex),
against goldens produced by the UNMODIFIED reference classes (tests/golden/sampler_variants_T64.npz, written by
oracle/make_golden_variants.py, which also pins the oracle restatement bit-exactly):

  reverse_diffusion + langevin   (correctors.py:37-64: batch-mean norms -> deterministic device reduction)
  reverse_diffusion + ald        (correctors.py:67-98, two inner steps)
  euler_maruyama   + none        (predictors.py:40-53 over RSDE.rsde_parts, sdes.py:128-150)
  condition="denoised" (+ sde_input "noisy" / "denoised"): the network conditioned on batch["fake"], the SDE anchored on
                                 the noisy or the denoised spectrogram (model_wrapper.py:281-299,321-328)
  condition="both" (the reference ScoreModel's default ctor): 6-channel network input cat[x, Y, Y_denoised]
                                 (model_wrapper.py:43-46,287-288), run as a 4 + 2 channel pyramid pair on the engine

Explicit noise in the reference's draw order (prior; per outer step the corrector's inner draws, then the predictor's).
Tolerances as for the default sampler: waveform rel-L2 <= 2e-3 (fp32 / TF32), <= 2e-2 (bf16).
"""
import os

import numpy as np
import pytest
import torch

import use_b200
from oracle import sgmse_oracle as O
from util import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-3, "bf16": 2e-2}
CASES = ["rd_langevin", "rd_ald", "em_none", "cond_denoised", "cond_denoised_sde_denoised", "cond_both_sde_denoised",
         "cond_both_sde_noisy"]


def _model(dtype, weight_seed, predictor="reverse_diffusion", corrector="none", condition="noisy", sde_input="noisy", **kw):
    m = use_b200.ScoreModel(backbone="ncsnpplarge", sde="ouve", t_eps=3e-2, condition=condition, sde_input=sde_input,
                            n_fft=1022, hop_length=160, num_frames=512, dtype=dtype, predictor=predictor,
                            corrector=corrector, **kw)
    m.score_net.load_state_dict(O.make_state_dict(O.LARGE6 if condition == "both" else O.LARGE, seed=weight_seed), strict=True)
    return m


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CASES)
def test_sampler_variant_matches_reference_golden(case, dtype):
    g = np.load(os.path.join(GOLDEN, "sampler_variants_T64.npz"))
    N, seed = int(g["N"]), int(g["seed"])
    pred, corr = str(g[f"{case}.predictor"]), str(g[f"{case}.corrector"])
    steps, snr = int(g[f"{case}.corrector_steps"]), float(g[f"{case}.snr"])
    condition, sde_input, key = str(g[f"{case}.condition"]), str(g[f"{case}.sde_input"]), str(g[f"{case}.key"])
    m = _model(dtype, int(g["weight_seed"]), pred, corr, condition=condition, sde_input=sde_input)
    y = torch.from_numpy(g["y"])
    per = O.draws_per_step(pred, corr, steps)
    noise = O.draw_noise((2, 1, 512, 64), N * per, seed).cuda()
    batch = {"perturbed": y.cuda()}
    if condition != "noisy" or sde_input != "noisy":
        batch["fake"] = torch.from_numpy(g["fake"]).cuda()  # the GAN stage's output (model_wrapper.py:271-272)
    got = m.sample(batch, N=N, corrector_steps=steps, snr=snr, noise=noise)[key].cpu()
    ref = torch.from_numpy(g[case])
    e = rel_l2(got, ref)
    assert bool(torch.isfinite(got).all()) and e <= TOL[dtype], (case, dtype, e)
    # a different variant is a different sampler: the default chain with the same leading draws must NOT match
    if case in ("rd_langevin", "rd_ald"):
        base = _model(dtype, int(g["weight_seed"])).sample({"perturbed": y.cuda()}, N=N, noise=noise[: N + 1].contiguous())
        assert rel_l2(base["enhanced"].cpu(), ref) > 10 * TOL[dtype]


def test_langevin_philox_is_deterministic_and_batch_coupled():
    """In-kernel Philox noise: same seed -> same bits (the norm reduction is a fixed-order two-pass sum, no float
    atomics); and the Langevin step size is a BATCH mean (correctors.py:55-57), so a clip sampled alone differs from
    the same clip inside a batch -- the reference's semantics, unlike the default sampler which is batch-invariant."""
    m = _model("bf16", 7, "reverse_diffusion", "langevin")
    y = O.synthetic_clips(2, 9600).cuda()
    a = m.sample({"perturbed": y}, N=2, snr=0.5, seed=3)["enhanced"]
    b = m.sample({"perturbed": y}, N=2, snr=0.5, seed=3)["enhanced"]
    assert torch.equal(a, b) and bool(torch.isfinite(a).all())
    alone = m.sample({"perturbed": y[:1]}, N=2, snr=0.5, seed=3)["enhanced"]
    assert not torch.equal(alone, a[:1])
    # ALD has per-sample step sizes: batch-invariant again
    m2 = _model("bf16", 7, "reverse_diffusion", "ald")
    a2 = m2.sample({"perturbed": y}, N=2, snr=0.4, seed=3)["enhanced"]
    alone2 = m2.sample({"perturbed": y[1:]}, N=2, snr=0.4, seed=3, clip0=1)["enhanced"]
    assert torch.equal(alone2, a2[1:])


def test_update_fn_single_steps_equal_the_fused_loop():
    """Driving the registry classes by hand (the reference's plugin usage) runs the same kernels: N update_fn calls of
    ReverseDiffusionPredictor with explicit zero noise reproduce the fused N-step chain's x_mean."""
    m = _model("fp32", 7)
    g = torch.Generator().manual_seed(21)
    B, F, T, N = 2, 512, 64, 3
    Y = (0.3 * torch.randn(B, 1, F, T, dtype=torch.complex64, generator=g)).cuda()
    noise = torch.zeros(N + 1, B, 1, F, T, dtype=torch.complex64, device="cuda")  # x0 = Y, no noise: deterministic
    fused, nfe = m.get_pc_sampler("reverse_diffusion", "none", Y, N=N, conditioning=[Y], noise=noise)()
    assert nfe == N
    sde = m.sde.copy()
    sde.N = N
    pred = use_b200.PredictorRegistry.get_by_name("reverse_diffusion")(sde, m)
    ts = sde.step_tables(N, m.t_eps)[0]
    x = Y
    for i in range(N):
        x_noisy, xm = pred.update_fn(x, torch.ones(B, device="cuda") * ts[i], Y, conditioning=[Y])
        x = xm  # follow the noise-free mean, as the zero-noise fused chain does
    # (not asserted bit-equal: the one-element schedule tables of a single step go through torch's scalar pow / log
    # paths on the host, which may differ from the vectorised ones by an ulp)
    assert rel_l2(torch.view_as_real(xm.cpu()), torch.view_as_real(fused.cpu())) < 1e-5
    # euler_maruyama and the correctors go through the same entry point
    em = use_b200.PredictorRegistry.get_by_name("euler_maruyama")(sde, m)
    xs, xm_em = em.update_fn(Y, torch.ones(B, device="cuda") * ts[0], Y, conditioning=[Y])
    assert xm_em.shape == Y.shape and rel_l2(torch.view_as_real(xm_em.cpu()), torch.view_as_real(fused.cpu())) < 1.0
    lc = use_b200.CorrectorRegistry.get_by_name("langevin")(sde, m, snr=0.5, n_steps=1)
    xs, xm_l = lc.update_fn(Y, torch.ones(B, device="cuda") * ts[0], Y, conditioning=[Y])
    assert bool(torch.isfinite(torch.view_as_real(xs)).all()) and not torch.equal(xs, xm_l)


def test_denoise_false_and_none_predictor():
    m = _model("bf16", 7)
    y = O.synthetic_clips(1, 9600).cuda()
    Y = m.stft_compressed(y).unsqueeze(1)
    noise = O.draw_noise((1, 1, 512, 64), 2, 5).cuda()
    mean, _ = m.get_pc_sampler("reverse_diffusion", "none", Y, N=2, conditioning=[Y], noise=noise)()
    state, _ = m.get_pc_sampler("reverse_diffusion", "none", Y, N=2, conditioning=[Y], noise=noise, denoise=False)()
    sde = m.sde.copy()
    sde.N = 2
    G_last = float(sde.step_tables(2, m.t_eps)[1][-1])
    # x = x_mean + G z of the last step
    assert rel_l2(torch.view_as_real((state - mean).cpu()), torch.view_as_real((G_last * noise[2]).cpu())) < 1e-3
    prior, nfe = m.get_pc_sampler("none", "none", Y, N=2, conditioning=[Y], noise=noise[:1].contiguous())()
    std1 = sde.step_tables(2, m.t_eps)[2]
    assert rel_l2(torch.view_as_real(prior.cpu()), torch.view_as_real((Y + std1 * noise[0]).cpu())) < 1e-6


def test_program_cache_survives_many_shapes():
    """the launch-program cache used to clear itself while a call still held a program pointer (two
    stream groups -> two lookups).  Walk more distinct (shape, workspace) combinations than the cache holds with B = 4
    (two groups), then re-run the first shape and compare bit for bit."""
    m = _model("bf16", 7)
    first = None
    for k, L in enumerate([9600, 19840, 30080, 9600 + 160 * 64 * 3, 50560, 60800, 71040, 9600]):
        y = O.synthetic_clips(4, L, seed=3).cuda()
        out = m.sample({"perturbed": y}, N=1, seed=11)["enhanced"]
        assert out.shape == (4, L) and bool(torch.isfinite(out).all())
        if k == 0:
            first = out.clone()
    assert torch.equal(out, first)


def test_minibatch_argument_keys_noise_by_global_clip():
    """get_pc_sampler(minibatch=k) must give every minibatch its own Philox streams (clip0 + offset)
    and its own slice of explicit noise: the split result equals the unsplit one bit for bit."""
    m = _model("bf16", 7)
    y = O.synthetic_clips(4, 9600).cuda()
    Y = m.stft_compressed(y).unsqueeze(1)
    whole, _ = m.get_pc_sampler("reverse_diffusion", "none", Y, N=2, conditioning=[Y], seed=9)()
    split, ns = m.get_pc_sampler("reverse_diffusion", "none", Y, N=2, minibatch=1, conditioning=[Y], seed=9)()
    assert torch.equal(whole, split) and ns == [2, 2, 2, 2]
    noise = O.draw_noise((4, 1, 512, 64), 2, 5).cuda()
    whole_n, _ = m.get_pc_sampler("reverse_diffusion", "none", Y, N=2, conditioning=[Y], noise=noise)()
    split_n, _ = m.get_pc_sampler("reverse_diffusion", "none", Y, N=2, minibatch=3, conditioning=[Y], noise=noise)()
    assert torch.equal(whole_n, split_n)


@pytest.mark.parametrize("dtype,tol", [("fp32", 5e-3), ("bf16", 3e-2)])
def test_ode_sampler_matches_reference_golden(dtype, tol):
    """sampler_type="ode" (sampling/__init__.py:76-159): scipy RK45 on the host around use_reverse_drift (one network
    evaluation + the fused probability-flow drift kernel per function evaluation), then the noise-free predictor step at
    t = eps.  Golden: the reference's own get_ode_sampler around an adapter score function (it raises with the
    reference's ScoreModel), rtol = atol = 1e-3.  The adaptive solver takes its step decisions from values that differ at
    the operand-rounding level, so the tolerance is the solver's own (a few 1e-3), not the network's."""
    g = np.load(os.path.join(GOLDEN, "sampler_variants_T64.npz"))
    N, seed = int(g["N"]), int(g["seed"])
    m = _model(dtype, int(g["weight_seed"]))
    y = torch.from_numpy(g["y"])
    noise = O.draw_noise((2, 1, 512, 64), 0, seed).cuda()  # the prior draw
    out = m.sample({"perturbed": y.cuda()}, sampler_type="ode", N=N, noise=noise,
                   ode_kwargs=dict(rtol=float(g["ode.rtol"]), atol=float(g["ode.atol"])))["enhanced"].cpu()
    e = rel_l2(out, torch.from_numpy(g["ode"]))
    assert bool(torch.isfinite(out).all()) and e <= tol, (dtype, e)
    # the drift entry point itself against the oracle's formula at one time
    Y = m.stft_compressed(y.cuda()).unsqueeze(1)
    x = Y + 0.2 * noise[0]
    sde = m.sde.copy()
    drift = m._reverse_drift(sde, x, Y, 0.4, [Y]).cpu()
    sd = O.make_state_dict(O.LARGE, seed=int(g["weight_seed"]))
    tt = torch.full((2,), 0.4)
    with torch.no_grad():
        score = -O.ncsnpp_forward(sd, O.LARGE, torch.cat([x.cpu(), Y.cpu()], 1), tt)
    gg = O.ouve_diffusion(tt)[:, None, None, None]
    ref = 1.5 * (Y.cpu() - x.cpu()) - gg**2 * score * 0.5
    assert rel_l2(torch.view_as_real(drift), torch.view_as_real(ref)) <= (5e-3 if dtype == "fp32" else 5e-2)
