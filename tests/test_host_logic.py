"""Host-side mirror of the reference interface: registries, config surface, state_dict compatibility, error
behaviour.  CPU only; nothing here launches a kernel."""
import os

import pytest
import torch

import use_b200
from use_b200.config import compose, instantiate
from use_b200.registry import Registry
from oracle import sgmse_oracle as O
from util import ROOT


def test_registry_contract():
    """register / get_by_name / get_all_names / double registration warning (util/registry.py:5-36)."""
    R = Registry("Thing")

    @R.register("a")
    class A:
        pass

    assert R.get_by_name("a") is A and R.get_all_names() == ["a"]
    with pytest.raises(ValueError, match="Thing with name 'zzz' unknown"):
        R.get_by_name("zzz")
    with pytest.warns(UserWarning, match="doubly registered"):
        @R.register("a")
        class B:
            pass
    assert R.get_by_name("a") is B


def test_registries_expose_reference_names():
    assert {"ncsnpp", "ncsnpplarge"} <= set(use_b200.BackboneRegistry.get_all_names())
    assert "ouve" in use_b200.SDERegistry.get_all_names()
    assert {"reverse_diffusion", "euler_maruyama", "none"} <= set(use_b200.PredictorRegistry.get_all_names())
    assert {"none", "langevin", "ald"} <= set(use_b200.CorrectorRegistry.get_all_names())


def test_state_dict_keys_match_reference_layout():
    """617 tensors / 64,799,782 parameters, identical names and shapes -> reference checkpoints load strict=True."""
    m = use_b200.ScoreModel(backbone="ncsnpplarge", condition="noisy", sde_input="noisy", n_fft=1022, hop_length=160)
    sd_ref = O.make_state_dict(O.LARGE)  # names/shapes validated against the reference by oracle/make_golden.py
    sd = m.score_net.state_dict()
    assert list(sd.keys()) != [] and set(sd.keys()) == set(sd_ref.keys())
    assert all(tuple(sd[k].shape) == tuple(sd_ref[k].shape) for k in sd)
    assert sum(v.numel() for v in sd.values()) == 64_799_782 and len(sd) == 617
    mod = use_b200.SGMSEModule(Score=m)
    assert all(k.startswith("Score.score_net.") for k in mod.state_dict().keys())
    mod.load_state_dict({"Score.score_net." + k: v for k, v in sd_ref.items()}, strict=True)
    assert torch.equal(m.score_net.all_modules[4].Conv_1.weight, sd_ref["all_modules.4.Conv_1.weight"])


def test_default_init_is_the_references_degenerate_one():
    """init_scale=0 -> 1e-10 variance scale for Conv_1 / NIN_3 / pyramid convs (layers.py:100-103)."""
    torch.manual_seed(0)
    net = use_b200.BackboneRegistry.get_by_name("ncsnpplarge")()
    assert float(net.all_modules[4].Conv_1.weight.std()) < 1e-6
    assert float(net.all_modules[4].Conv_0.weight.std()) > 1e-3
    assert float(net.all_modules[31].NIN_3.W.std()) < 1e-6
    assert float(net.all_modules[73].weight.std()) < 1e-6
    assert abs(float(net.all_modules[0].W.std()) - 16) < 4


def test_config_surface_composes_and_instantiates():
    cfg = compose(os.path.join(ROOT, "configs"), "predict.yaml", ["model.Score.N=30", "model.Score.dtype=bf16"])
    assert cfg["model"]["_target_"] == "src.models.SGMSE_module.SGMSEModule"
    assert cfg["model"]["Score"]["t_eps"] == pytest.approx(3e-2) and cfg["model"]["optimizer"]["lr"] == pytest.approx(5e-4)
    m = instantiate(cfg["model"])
    assert isinstance(m, use_b200.SGMSEModule) and isinstance(m.Score, use_b200.ScoreModel)
    assert m.Score.default_N == 30 and m.Score.dtype_name == "bf16" and m.Score.n_fft == 1022
    assert callable(m.optimizer)  # _partial_: true
    # the reference's _target_ strings also resolve through the src/ shim package
    from src.models.components.sgmse.model_wrapper import ScoreModel as Shim
    assert Shim is use_b200.ScoreModel


def test_unsupported_options_fail_loudly():
    with pytest.raises(NotImplementedError):
        use_b200.ScoreModel(backbone="ncsnpplarge", condition="neither", sde_input="denoised")
    both = use_b200.ScoreModel(backbone="ncsnpplarge")  # the reference's default ctor: condition="both" -> 6 input channels
    assert both.score_net.input_channels == 6 and tuple(both.score_net.all_modules[3].weight.shape) == (128, 6, 3, 3)
    sd6 = O.make_state_dict(O.LARGE6)  # names / shapes validated against the reference by oracle/make_golden_variants.py
    assert {k: tuple(v.shape) for k, v in both.score_net.state_dict().items()} == {k: tuple(v.shape) for k, v in sd6.items()}
    with pytest.raises(ValueError, match="unknown"):
        use_b200.ScoreModel(backbone="does_not_exist", condition="noisy", sde_input="noisy")
    with pytest.raises(NotImplementedError):
        use_b200.NCSNpp(resblock_type="ddpm")
    m = use_b200.ScoreModel(backbone="ncsnpplarge", condition="noisy", sde_input="noisy", n_fft=1022, hop_length=160)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.sample({"perturbed": torch.zeros(1, 9600)}, N=1)  # CPU tensors: no CPU path, no silent fallback
    with pytest.raises(NotImplementedError):  # the ODE sampler needs the B200 model's drift entry point, not any callable
        from use_b200 import sampling
        sampling.get_ode_sampler(m.sde.copy(), lambda x, t, y: x, torch.zeros(1, 1, 8, 8, dtype=torch.complex64))
    with pytest.raises(NotImplementedError):
        m.sample({"perturbed": torch.zeros(1, 9600)}, sampler_type="bogus")


def test_plugin_predictor_corrector_run_through_the_host_loop():
    """Third-party predictors / correctors registered through the plugin API (no fused `kind`) are sequenced by the host
    loop -- corrector then predictor per step, over linspace(T, eps, N) -- around any score function; ReverseSDE gives
    them the reverse drift.  CPU stand-in score function: checks the API contract, not the kernels."""
    from use_b200 import sampling

    sde = use_b200.OUVESDE()
    sde.N = 4
    y = torch.randn(2, 1, 8, 8, dtype=torch.complex64)
    calls = []

    def score_fn(x, t, score_conditioning=None, sde_input=None):
        assert score_conditioning[0] is y and sde_input is y
        return -(x - y)

    if "test_plugin_pred" not in use_b200.PredictorRegistry.get_all_names():
        @use_b200.PredictorRegistry.register("test_plugin_pred")
        class PluginPredictor(sampling.Predictor):
            def update_fn(self, x, t, y_, conditioning=None):
                calls.append(("p", float(t[0])))
                f, G = self.rsde.discretize(x, t, y_, conditioning=conditioning)
                return x - f, x - f  # noise-free reverse-diffusion step

        @use_b200.CorrectorRegistry.register("test_plugin_corr")
        class PluginCorrector(sampling.Corrector):
            def update_fn(self, x, t, y_, conditioning=None):
                calls.append(("c", float(t[0])))
                drift, g = self.rsde.sde(x, t, y_, conditioning=conditioning)
                assert drift.shape == x.shape and g.shape == (2, 1, 1, 1)
                return x, x

    s = use_b200.get_pc_sampler("test_plugin_pred", "test_plugin_corr", sde, score_fn, y, corrector_steps=1, snr=0.5,
                                conditioning=[y])
    x, n = s()
    assert x.shape == y.shape and bool(torch.isfinite(torch.view_as_real(x)).all()) and n == 8
    ts = [float(v) for v in torch.linspace(1, 3e-2, 4)]
    assert calls == [(k, t) for t in ts for k in ("c", "p")]
    # a built-in (fused) predictor refuses a score function that is not the B200 ScoreModel: there is no host fallback
    with pytest.raises(NotImplementedError):
        use_b200.get_pc_sampler("reverse_diffusion", "test_plugin_corr", sde, score_fn, y, conditioning=[y])()
    # NonePredictor / NoneCorrector are pure pass-throughs
    x0, n0 = use_b200.get_pc_sampler("none", "none", sde, score_fn, y, conditioning=[y])()
    assert x0.shape == y.shape and n0 == 4


def test_variant_tables_match_reference_expressions():
    """g(t_i) and the ALD step sizes handed to the fused loop = the reference's float32 torch expressions."""
    sde = use_b200.OUVESDE()
    ts, G, std1 = sde.step_tables(30, 3e-2)
    g, ald = sde.variant_tables(ts, 0.4)
    og = O.ouve_diffusion(ts)
    assert torch.equal(g, og.to(torch.float32))
    assert torch.equal(ald, (0.4 * O.ouve_std(ts)) ** 2 * 2)
    assert torch.equal(G, O.step_coefficients(30)[1])


def test_gan_generator_host_structure():
    """NCSNPP_Wrapper / GANModule (LSGAN stage): discriminative NCSN++ parameter layout and config surface."""
    G = use_b200.NCSNPP_Wrapper(n_fft=1022, hop_length=160, num_frames=480)
    sd_ref = O.make_state_dict(O.GAN_G, seed=13)  # names / shapes validated against the reference by make_golden.py
    sd = G.net.state_dict()
    assert set(sd.keys()) == set(sd_ref.keys())
    assert all(tuple(sd[k].shape) == tuple(sd_ref[k].shape) for k in sd)
    G.net.load_state_dict(sd_ref, strict=True)
    cfg = compose(os.path.join(ROOT, "configs"), "predict.yaml", ["model=LSGAN"])
    m = instantiate(cfg["model"])
    assert isinstance(m, use_b200.GANModule) and isinstance(m.G, use_b200.NCSNPP_Wrapper) and m.G.n_fft == 1022
    m.load_state_dict({"G.net." + k: v for k, v in sd_ref.items()})
    assert torch.equal(m.G.net.all_modules[1].weight, sd_ref["all_modules.1.weight"])
    with pytest.raises(NotImplementedError):
        G({"clean": torch.zeros(1, 8), "perturbed": torch.zeros(1, 8)})
    with pytest.raises(RuntimeError, match="CUDA"):
        G({"perturbed": torch.zeros(1, 9600)})


def test_loadwav_datamodule_file_discovery(tmp_path):
    """The reference's four ways to name the inputs (loadwav_dataset.py:38-77): json lines, list file, in-memory lists,
    folder walk -- host logic only."""
    import json

    from use_b200.predict import LoadWavDataModule

    (tmp_path / "a").mkdir()
    for n in ("x.wav", "y.wav", "z.txt"):
        (tmp_path / "a" / n).write_bytes(b"")
    walk = LoadWavDataModule(data_folder=str(tmp_path))
    assert sorted(os.path.basename(f) for f in walk.files) == ["x.wav", "y.wav"]
    lst = tmp_path / "list.txt"
    lst.write_text("/p/1.wav\n\n/p/2.wav\n")
    assert LoadWavDataModule(list_path=str(lst)).files == ["/p/1.wav", "/p/2.wav"]
    js = tmp_path / "m.json"
    js.write_text(json.dumps({"audio_filepath": "/q/1.wav"}) + "\n" + json.dumps({"file_path": "/q/2.wav"}) + "\n"
                  + json.dumps({"audio_filepath": "/q/1.wav"}) + "\n")
    assert LoadWavDataModule(json_path=str(js)).files == ["/q/1.wav", "/q/2.wav"]
    assert LoadWavDataModule(input_plain_list=["/r/1.wav"]).files == ["/r/1.wav"]
    assert LoadWavDataModule(input_json_list=[json.dumps({"file_path": "/s.wav"})]).files == ["/s.wav"]
    with pytest.raises(ValueError, match="No input list"):
        LoadWavDataModule()
    with pytest.raises(NotImplementedError):
        LoadWavDataModule(data_folder=str(tmp_path), output_resample=True)


@pytest.mark.parametrize("model_cfg", ["SGMSE_Large", "LSGAN"])
def test_load_checkpoint_for_both_module_kinds(tmp_path, model_cfg):
    """predict() calls model.load_checkpoint(ckpt_path) for whatever `model=` selects (GANModule had none):
    a Lightning-style .ckpt ({"state_dict": ...}) round-trips through both module classes."""
    cfg = compose(os.path.join(ROOT, "configs"), "predict.yaml", [f"model={model_cfg}"])
    m = instantiate(cfg["model"])
    sd = {k: torch.full_like(v, 0.25) if v.is_floating_point() else v for k, v in m.state_dict().items()}
    path = str(tmp_path / "fake.ckpt")
    torch.save({"state_dict": sd, "epoch": 3}, path)
    m.load_checkpoint(path)
    got = m.state_dict()
    k = next(k for k in got if k.endswith("weight"))
    assert float(got[k].flatten()[0]) == 0.25 and set(got.keys()) == set(sd.keys())


class _FakeEngine:
    """Records what the sampler asks of the engine (no kernels): latency-mode pinning and per-call batch sizes."""

    def __init__(self, fail_on_call=None):
        self.modes, self.calls, self.fail_on_call = [], [], fail_on_call

    def latency_mode(self, job_clips):
        self.modes.append(job_clips)

    def pc_sample(self, Y, ts, G, std1, **kw):
        self.calls.append((int(Y.shape[0]), int(kw["clip0"])))
        if self.fail_on_call is not None and len(self.calls) == self.fail_on_call:
            raise RuntimeError("boom")
        return torch.zeros_like(Y), torch.zeros_like(Y)


def _fake_model(monkeypatch, eng, **kw):
    m = use_b200.ScoreModel(backbone="ncsnpplarge", condition="noisy", sde_input="noisy", n_fft=1022, hop_length=160, **kw)
    monkeypatch.setattr(m, "_engine", lambda device: eng)
    return m


def test_latency_mode_is_pinned_from_the_whole_job(monkeypatch):
    """One job = one kernel mode (include/use_b200.h "ksplit"): micro-batches, minibatches and shards of a job pin the
    engine's latency mode from the size of the WHOLE job and release it afterwards, also when a call fails."""
    Y = torch.zeros(5, 1, 512, 64, dtype=torch.complex64)
    eng = _FakeEngine()
    m = _fake_model(monkeypatch, eng, micro_batch=2)
    m.get_pc_sampler("reverse_diffusion", "none", Y, N=2, conditioning=[Y], seed=1)()
    assert eng.modes == [5, None] and eng.calls == [(2, 0), (2, 2), (1, 4)]   # 2 + 2 + 1 clips, one mode (5 clips)
    # a shard of a larger job names the job's size; a job of two clips runs in latency mode by its own size
    eng.modes.clear(); eng.calls.clear()
    m.micro_batch = None
    m.get_pc_sampler("reverse_diffusion", "none", Y[:1], N=2, conditioning=[Y[:1]], seed=1, clip0=3, job_clips=4)()
    m.get_pc_sampler("reverse_diffusion", "none", Y[:2], N=2, conditioning=[Y[:2]], seed=1)()
    assert eng.modes == [4, None, 2, None] and eng.calls == [(1, 3), (2, 0)]
    # the reference's minibatch argument: every minibatch carries the whole batch as its job
    eng.modes.clear(); eng.calls.clear()
    m.get_pc_sampler("reverse_diffusion", "none", Y, N=2, minibatch=2, conditioning=[Y], seed=1)()
    assert eng.modes == [5, None, 5, None, 5, None] and eng.calls == [(2, 0), (2, 2), (1, 4)]
    # released in a finally block
    bad = _FakeEngine(fail_on_call=1)
    m2 = _fake_model(monkeypatch, bad)
    with pytest.raises(RuntimeError, match="boom"):
        m2.get_pc_sampler("reverse_diffusion", "none", Y, N=2, conditioning=[Y], seed=1)()
    assert bad.modes == [5, None]
