"""The C-ABI library: loads without a GPU, exports exactly what include/use_b200.h declares, and its host-side
logic (architecture walk, weight validation/packing, workspace planning) works on CPU.  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from use_b200 import _lib
from oracle import sgmse_oracle as O
from util import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "use_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(use_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_python_symbol_table():
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (use_[a-z0-9_]+)", out))
    assert exported == set(header_symbols())
    for name in header_symbols():
        assert hasattr(L, name)
    assert L.use_abi_version() == 5


def _engine(L, net, dt):
    cfg = _lib.UseConfig()
    cfg.nf, cfg.num_levels, cfg.num_res_blocks, cfg.input_channels, cfg.act_dtype = net.nf, len(net.ch_mult), net.num_res_blocks, 4, dt
    for i, m in enumerate(net.ch_mult):
        cfg.ch_mult[i] = m
    cfg.n_fft, cfg.hop, cfg.spec_factor, cfg.spec_abs_exponent, cfg.theta = 1022, 160, 0.15, 0.5, 1.5
    cfg.conditional, cfg.scale_by_sigma = int(net.conditional), int(net.scale_by_sigma)
    cfg.input_channels = net.input_channels
    return L.use_engine_create(C.byref(cfg))


def _feed(L, h, sd):
    for k, v in sd.items():
        v = v.contiguous()
        shp = (C.c_int64 * max(v.dim(), 1))(*v.shape)
        assert L.use_engine_set_weight(h, k.encode(), v.data_ptr(), shp, v.dim()) == 0


@pytest.mark.parametrize("dt", [0, 1])
def test_engine_host_logic_tiny(dt):
    L = _lib.lib()
    h = _engine(L, O.TINY, dt)
    assert h, L.use_last_error()
    sd = O.make_state_dict(O.TINY, seed=11)
    _feed(L, h, sd)
    n = C.c_size_t()
    assert L.use_engine_pack(h, C.byref(n)) == 0, L.use_last_error()
    assert n.value > sum(v.numel() for v in sd.values())  # at least 1 byte per parameter
    w = C.c_size_t()
    assert L.use_engine_workspace_bytes(h, 2, 16, 24, C.byref(w)) == 0, L.use_last_error()
    w2 = C.c_size_t()
    assert L.use_engine_workspace_bytes(h, 4, 16, 24, C.byref(w2)) == 0
    assert w2.value > w.value > 0
    # error behaviour: size not divisible by 2^(levels-1), like pad_spec's contract
    assert L.use_engine_workspace_bytes(h, 1, 16, 25, C.byref(w)) != 0
    assert b"divisible" in L.use_last_error()
    L.use_engine_destroy(h)


def test_engine_host_logic_discriminative_generator():
    """The LSGAN generator configuration (2 input channels, no time embedding) plans and packs on CPU."""
    L = _lib.lib()
    h = _engine(L, O.GAN_TINY, 1)
    assert h, L.use_last_error()
    _feed(L, h, O.make_state_dict(O.GAN_TINY, seed=3))
    n, w = C.c_size_t(), C.c_size_t()
    assert L.use_engine_pack(h, C.byref(n)) == 0, L.use_last_error()
    assert L.use_engine_workspace_bytes(h, 2, 16, 24, C.byref(w)) == 0 and w.value > 0, L.use_last_error()
    L.use_engine_destroy(h)


def test_engine_options():
    """A/B switches of the engine: known keys are accepted (and both plans still fit their workspace), unknown ones and
    out-of-range values fail with a message."""
    L = _lib.lib()
    h = _engine(L, O.TINY, 1)
    _feed(L, h, O.make_state_dict(O.TINY, seed=11))
    assert L.use_engine_pack(h, None) == 0, L.use_last_error()
    sizes = {}
    for fuse in (1, 0):
        assert L.use_engine_set_option(h, b"fuse_gn", fuse) == 0
        w = C.c_size_t()
        assert L.use_engine_workspace_bytes(h, 2, 16, 24, C.byref(w)) == 0, L.use_last_error()
        sizes[fuse] = w.value
    assert sizes[0] > 0 and sizes[1] > 0
    for key, val in ((b"use_graphs", 0), (b"use_graphs", 1), (b"overlap_groups", 1), (b"overlap_groups", 2), (b"fuse_head", 0), (b"fuse_head", 1)):
        assert L.use_engine_set_option(h, key, val) == 0, L.use_last_error()
    assert L.use_engine_set_option(h, b"overlap_groups", 3) != 0 and b"overlap_groups" in L.use_last_error()
    assert L.use_engine_set_option(h, b"no_such_option", 1) != 0 and b"unknown option" in L.use_last_error()
    L.use_engine_destroy(h)


def test_engine_rejects_missing_and_misshaped_weights():
    L = _lib.lib()
    h = _engine(L, O.TINY, 0)
    sd = O.make_state_dict(O.TINY, seed=11)
    missing = dict(sd)
    missing.pop("all_modules.4.Conv_1.weight")
    _feed(L, h, missing)
    assert L.use_engine_pack(h, None) != 0
    assert b"all_modules.4.Conv_1.weight" in L.use_last_error()
    bad = torch.zeros(3, 3)
    shp = (C.c_int64 * 2)(3, 3)
    L.use_engine_set_weight(h, b"all_modules.4.Conv_1.weight", bad.data_ptr(), shp, 2)
    assert L.use_engine_pack(h, None) != 0 and b"shape" in L.use_last_error()
    L.use_engine_destroy(h)


def test_unsupported_architecture_fails_loudly():
    L = _lib.lib()
    cfg = _lib.UseConfig()
    cfg.nf, cfg.num_levels, cfg.num_res_blocks, cfg.input_channels, cfg.act_dtype = 96, 2, 1, 4, 1  # ncsnpp12M-like width
    cfg.ch_mult[0], cfg.ch_mult[1] = 1, 2
    cfg.conditional, cfg.scale_by_sigma = 1, 1
    assert not L.use_engine_create(C.byref(cfg))
    assert b"not supported" in L.use_last_error()


def test_pack_conv_weight_layout():
    L = _lib.lib()
    w = torch.arange(2 * 3 * 9, dtype=torch.float32).reshape(2, 3, 3, 3)
    out = torch.empty(9 * 2 * 3, dtype=torch.float32)
    assert L.use_pack_conv_weight(0, w.data_ptr(), 2, 3, 3, out.data_ptr()) == 0
    assert torch.equal(out.reshape(9, 2, 3), w.reshape(2, 3, 9).permute(2, 0, 1))  # [tap][O][I]
    outb = torch.empty(9 * 2 * 3, dtype=torch.bfloat16)
    assert L.use_pack_conv_weight(1, w.data_ptr(), 2, 3, 3, outb.data_ptr()) == 0
    assert torch.equal(outb.reshape(9, 2, 3), w.reshape(2, 3, 9).permute(2, 0, 1).to(torch.bfloat16))


def test_pack_head_weight_layout():
    """Pyramid-head weights: [pc][C][3][3] -> [48][C] with row tap * pc + c_out, zero rows above 9 * pc."""
    L = _lib.lib()
    for pc in (4, 2):
        Cc = 64
        w = torch.randn(pc, Cc, 3, 3)
        out = torch.full((48, Cc), float("nan"), dtype=torch.float32)
        assert L.use_pack_head_weight(0, w.contiguous().data_ptr(), pc, Cc, out.data_ptr()) == 0
        ref = torch.zeros(48, Cc)
        for tap in range(9):
            for co in range(pc):
                ref[tap * pc + co] = w[co, :, tap // 3, tap % 3]
        # fp32 mode stores TF32-rounded values
        from util import round_tf32
        assert torch.equal(out, round_tf32(ref))
        outb = torch.empty(48, Cc, dtype=torch.bfloat16)
        assert L.use_pack_head_weight(1, w.contiguous().data_ptr(), pc, Cc, outb.data_ptr()) == 0
        assert torch.equal(outb, ref.to(torch.bfloat16))
    assert L.use_pack_head_weight(0, w.data_ptr(), 3, 64, out.data_ptr()) != 0


def test_ncu_summary_tool_reproduces_committed_summary(tmp_path):
    """tools/summarize_ncu.py on the committed launch list gives the committed per-kernel summary (the file bench.py
    reads `roofline.traffic` from)."""
    import json
    import sys
    for d in ("fp32", "bf16"):
        src = os.path.join(ROOT, "profiles", f"r01_ncu_launches_{d}_b2.csv")
        out = tmp_path / f"s_{d}.json"
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_ncu.py"), src, str(out)], check=True,
                       capture_output=True)
        got = json.load(open(out))["kernels"]
        ref = json.load(open(os.path.join(ROOT, "profiles", f"r01_kernel_summary_{d}.json")))["kernels"]
        assert got == ref
        conv = [v for k, v in got.items() if "conv_tc_kernel" in k]
        assert sum(v["launches"] for v in conv) == 99  # the convolutions of one NCSNppLarge evaluation


def test_missing_library_is_a_loud_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no fallback"):
        _lib.lib()
