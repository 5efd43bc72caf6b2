"""Goldens for the other registered sampler variants, from the UNMODIFIED reference.

TEST INFRASTRUCTURE (build container only: needs /root/reference).

  1. reverse_diffusion + langevin   ScoreModel(corrector="langevin").sample(batch, N=3, snr=0.5)        (correctors.py:37-64)
  2. reverse_diffusion + ald        ScoreModel(corrector="ald").sample(batch, N=3, corrector_steps=2, snr=0.4)  (:67-98)
  3. euler_maruyama + none          the reference's own sampling.get_pc_sampler + EulerMaruyamaPredictor
     (predictors.py:40-53) around an ADAPTER score function: with the reference's ScoreModel as score_fn the call
     raises TypeError (RSDE.rsde_parts passes (x, t, conditioning) / (x, t, y), sdes.py:131-134, while
     ScoreModel.forward needs sde_input too), so the predictor is driven with ``lambda x, t, y: m(x, t, [y], y)``.
     Predictor, reverse SDE and sampler loop are the reference's classes, untouched.

Each case: NCSNppLarge with the oracle's seeded weights, B=2, 0.4 s clips (T=64), torch.manual_seed(42); asserts that
the oracle restatement (oracle/sgmse_oracle.py: pc_sample_spec) is BIT-exact and writes
tests/golden/sampler_variants_T64.npz.

    python oracle/make_golden_variants.py
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

CASES = {
    "rd_langevin": dict(predictor="reverse_diffusion", corrector="langevin", corrector_steps=1, snr=0.5),
    "rd_ald": dict(predictor="reverse_diffusion", corrector="ald", corrector_steps=2, snr=0.4),
    "em_none": dict(predictor="euler_maruyama", corrector="none", corrector_steps=1, snr=0.5),
    # the GAN-refiner conditioning variants (batch["fake"] = the first stage's output; model_wrapper.py:281-299,321-328)
    "cond_denoised": dict(predictor="reverse_diffusion", corrector="none", corrector_steps=1, snr=0.5),
    "cond_denoised_sde_denoised": dict(predictor="reverse_diffusion", corrector="none", corrector_steps=1, snr=0.5),
    # the reference ScoreModel's DEFAULT ctor: condition="both" (6-channel network input), sde_input="denoised"
    "cond_both_sde_denoised": dict(predictor="reverse_diffusion", corrector="none", corrector_steps=1, snr=0.5),
    "cond_both_sde_noisy": dict(predictor="reverse_diffusion", corrector="none", corrector_steps=1, snr=0.5),
    # probability-flow ODE sampler (sampling/__init__.py:76-159): scipy RK45 on the host, rtol = atol = 1e-3 to bound the
    # number of function evaluations; like euler_maruyama it only runs in the reference around an adapter score function
    "ode": dict(sampler_type="ode", rtol=1e-3, atol=1e-3),
}
COND = {"cond_denoised": ("denoised", "noisy", "enhanced"),
        "cond_denoised_sde_denoised": ("denoised", "denoised", "fake_sde_enhanced"),
        "cond_both_sde_denoised": ("both", "denoised", "fake_sde_enhanced"),
        "cond_both_sde_noisy": ("both", "noisy", "enhanced")}


def main():
    torch.set_num_threads(os.cpu_count())
    ScoreModel, _, _, _ = import_reference()
    from src.models.components.sgmse import sampling as RS  # type: ignore

    sd4, sd6 = O.make_state_dict(O.LARGE, seed=7), O.make_state_dict(O.LARGE6, seed=7)
    B, L, N, seed = 2, 9600, 3, 42
    y = O.synthetic_clips(B, L)
    fake = 0.7 * y + 0.05 * O.synthetic_clips(B, L, seed=77)  # stand-in for the GAN stage's denoised output
    out = dict(y=y.numpy(), fake=fake.numpy(), B=B, L=L, N=N, seed=seed, weight_seed=7)
    for name, kw in CASES.items():
        condition, sde_input, key = COND.get(name, ("noisy", "noisy", "enhanced"))
        m = ScoreModel(backbone="ncsnpplarge", sde="ouve", t_eps=3e-2, mode="regen-joint-training", condition=condition,
                       loss_type="mse", n_fft=1022, hop_length=160, num_frames=512, window="hann", spec_factor=0.15,
                       spec_abs_exponent=0.5, sde_input=sde_input, predictor=kw.get("predictor", "reverse_diffusion"),
                       corrector=kw.get("corrector", "none")).eval()
        sdL, net = (sd6, O.LARGE6) if condition == "both" else (sd4, O.LARGE)
        m.score_net.load_state_dict(sdL, strict=True)
        torch.manual_seed(seed)
        if name == "em_none":
            with torch.no_grad():
                Y = O.pad_spec(m.spec_fwd(m.stft(y)).unsqueeze(1))
                sde = m.sde.copy()
                sde.N = N
                sampler = RS.get_pc_sampler("euler_maruyama", "none", sde=sde, score_fn=lambda x, t, yy: m(x, t, [yy], yy),
                                            y=Y, eps=m.t_eps, conditioning=None)
                xm, nfe = sampler()
                ref = m.istft(m.spec_back(xm.squeeze(1)), L)
        elif name == "ode":
            with torch.no_grad():
                Y = O.pad_spec(m.spec_fwd(m.stft(y)).unsqueeze(1))
                sde = m.sde.copy()
                sde.N = N
                sampler = RS.get_ode_sampler(sde, lambda x, t, yy: m(x, t, [yy], yy), y=Y, eps=m.t_eps, rtol=kw["rtol"],
                                             atol=kw["atol"], device="cpu")
                xm, nfe = sampler()
                print("ode nfe =", nfe, flush=True)
                out["ode.nfe"] = nfe
                ref = m.istft(m.spec_back(xm.squeeze(1)), L)
        elif name in COND:
            ref = m.sample({"perturbed": y.clone(), "fake": fake.clone()}, N=N)[key]
        else:
            ref = m.sample({"perturbed": y.clone()}, N=N, corrector_steps=kw["corrector_steps"], snr=kw["snr"])["enhanced"]
        mine = O.sample(sdL, y, N, seed=seed, net=net, fake=fake if name in COND else None, condition=condition,
                        sde_input=sde_input, **kw)
        d = float((ref - mine).abs().max())
        print(f"{name}: max|ref-oracle| = {d}", flush=True)
        assert d == 0.0, name
        out[name] = ref.numpy()
        for k, v in kw.items():
            out[f"{name}.{k}"] = v
        out[f"{name}.condition"], out[f"{name}.sde_input"], out[f"{name}.key"] = condition, sde_input, key
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sampler_variants_T64.npz"), **out)
    print("written tests/golden/sampler_variants_T64.npz")


if __name__ == "__main__":
    main()
