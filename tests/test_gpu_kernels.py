"""Per-kernel parity on the B200, called through the C ABI (libuse_b200.so); torch CPU float64 is the checker.

Operands are rounded to what the tensor core consumes (bf16 / TF32) BEFORE the reference computation, so the
remaining difference is fp32 accumulation order (+ the output rounding of bf16 storage): tolerances are tight
enough that any layout / descriptor / swizzle / pipeline-phase bug shows up as O(1) errors.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as Fnn

from use_b200 import _lib
from oracle import sgmse_oracle as O
from util import (BF16, F32, act_tensor, describe_mismatch, from_act, int_array, pack_weight, ptr_array, rel_l2, stream,
                  to_operand)

pytestmark = pytest.mark.gpu

OUT_TOL = {F32: 2e-5, BF16: 6e-3}  # relative-to-max tolerance on act-dtype outputs


def _sync():
    torch.cuda.synchronize()


def run_conv_tc(dt, segs, B, H, W, N, bias, res=None, scale=1.0, want_stats=False):
    """segs: list of (x_nchw fp32 already operand-rounded, w_oihw fp32 already rounded, ks).  One weight tensor per
    segment here (wc0 = 0); the concat-window form is covered by test_conv_tc_weight_window."""
    L = _lib.lib()
    acts, ws, ct, c0, cc, cw, wc0, taps, keep = [], [], [], [], [], [], [], [], []
    for x, w, ks in segs:
        a = act_tensor(x, dt)
        pw = pack_weight(L, w, dt)
        keep += [a, pw]
        acts.append(a.data_ptr()); ws.append(pw.data_ptr())
        ct.append(x.shape[1]); c0.append(0); cc.append(x.shape[1]); cw.append(w.shape[1]); wc0.append(0); taps.append(ks * ks)
    out = torch.empty(B, H, W, N, device="cuda", dtype=torch.bfloat16 if dt == BF16 else torch.float32)
    bias_d = bias.to("cuda", torch.float32).contiguous()
    bstride = N if bias.dim() == 2 else 0
    res_d = act_tensor(res, dt) if res is not None else None
    stats = None
    if want_stats:
        stats = torch.zeros(B, N, 2, dtype=torch.int64, device="cuda")
    rc = L.use_op_conv_tc(dt, len(segs), ptr_array(acts), int_array(ct), int_array(c0), int_array(cc), ptr_array(ws),
                          int_array(cw), int_array(wc0), int_array(taps), B, H, W, N, bias_d.data_ptr(), bstride,
                          res_d.data_ptr() if res_d is not None else None, float(scale), out.data_ptr(),
                          stats.data_ptr() if want_stats else None, stream())
    assert rc == 0, L.use_last_error()
    _sync()
    if want_stats:
        return from_act(out), stats_to_float(stats)
    return from_act(out)


def ref_conv(dt, segs, bias, res=None, scale=1.0):
    acc = None
    for x, w, ks in segs:
        y = Fnn.conv2d(x.double(), w.double(), padding=ks // 2)
        acc = y if acc is None else acc + y
    acc = acc + (bias.double()[:, :, None, None] if bias.dim() == 2 else bias.double()[None, :, None, None])
    if res is not None:
        acc = acc + to_operand(res, dt).double() if dt == BF16 else acc + res.double()
    return (acc * scale).float()


def make_seg(g, dt, B, Cin, Cout, H, W, ks):
    x = to_operand(torch.randn(B, Cin, H, W, generator=g), dt)
    w = to_operand(torch.randn(Cout, Cin, ks, ks, generator=g) / np.sqrt(Cin * ks * ks), dt)
    return (x, w, ks)


CONV_CASES = [
    # name, B, H, W, N, [(Cin, ks), ...], per-sample bias, residual
    ("1x1_single_chunk", 1, 32, 8, 128, [(64, 1)], False, False),
    ("1x1_multi_chunk", 1, 32, 8, 128, [(256, 1)], False, False),
    ("3x3_one_tile", 1, 32, 8, 128, [(64, 3)], False, False),
    ("3x3_c128", 1, 32, 16, 128, [(128, 3)], False, False),
    ("3x3_ragged_edges", 2, 20, 10, 128, [(128, 3)], True, False),
    ("3x3_n256", 2, 16, 24, 256, [(128, 3)], True, False),
    ("3x3_n256_small_level", 1, 8, 10, 256, [(256, 3)], False, True),
    ("3x3_n64", 1, 16, 24, 64, [(64, 3)], False, False),
    ("resblock_tail_3seg", 1, 32, 16, 128, [(128, 3), (128, 1), (64, 1)], False, False),
    ("residual_scale", 2, 32, 8, 128, [(128, 3)], True, True),
    ("persistent_many_tiles", 2, 128, 160, 128, [(128, 3)], False, True),
    ("persistent_many_tiles_n256", 3, 64, 80, 256, [(256, 3), (256, 1)], True, False),
]


@pytest.mark.parametrize("dt", [F32, BF16], ids=["tf32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_tc(case, dt):
    name, B, H, W, N, seg_spec, per_sample, with_res = case
    g = torch.Generator().manual_seed(hash(name) % 1000)
    segs = [make_seg(g, dt, B, cin, N, H, W, ks) for cin, ks in seg_spec]
    bias = torch.randn(B, N, generator=g) if per_sample else torch.randn(N, generator=g)
    res = torch.randn(B, N, H, W, generator=g) if with_res else None
    scale = 0.70710678 if (with_res or len(segs) > 1) else 1.0
    got = run_conv_tc(dt, segs, B, H, W, N, bias, res, scale)
    ref = ref_conv(dt, segs, bias, res, scale)
    tol = OUT_TOL[dt] * float(ref.abs().max())
    assert float((got - ref).abs().max()) <= tol, f"{name}: " + describe_mismatch(got, ref)


KS_CASES = [
    # name, B, H, W, [(Cin, ks), ...], per-sample bias, residual   (C_out = 256: the only family with a split-K form)
    ("ks4_bottleneck_8x10", 3, 8, 10, [(256, 3)], True, True),            # 1 tile per clip -> clusters of 4
    ("ks4_16x20_cat", 2, 16, 20, [(256, 3), (256, 3)], True, False),      # 3 tiles, two 3x3 segments (concatenated input)
    ("ks4_16x20_skip", 2, 16, 20, [(256, 3), (256, 1), (256, 1)], False, False),  # 3x3 + two 1x1 skip segments
    ("ks2_32x40", 2, 32, 40, [(256, 3)], True, True),                     # 10 tiles -> clusters of 2
    ("ks2_32x24_ragged", 3, 30, 20, [(128, 3), (256, 1)], False, False),  # ragged edges, odd tap split
    ("ks4_one_chunk", 1, 8, 8, [(64, 3)], False, False),                  # 9 taps over 4 ranks (fp32: 18 over 4)
]


@pytest.mark.parametrize("dt", [F32, BF16], ids=["tf32", "bf16"])
@pytest.mark.parametrize("case", KS_CASES, ids=[c[0] for c in KS_CASES])
def test_conv_tc_latency_split_k(case, dt):
    """conv_tc_ks.cuh: the split-K cluster form of the low-resolution levels (latency-mode programs): clusters of 2 / 4
    CTAs accumulate disjoint tap ranges and the leader adds the partial sums in rank order through DSMEM.  vs fp64 torch,
    vs the single-accumulator kernel (same operands, another association: fp32 rounding only), statistics of the output,
    and batch invariance INSIDE the mode (a clip alone == the same clip in a batch, bit for bit)."""
    name, B, H, W, seg_spec, per_sample, with_res = case
    N = 256
    L = _lib.lib()
    g = torch.Generator().manual_seed(len(name) * 7 + H)
    segs = [make_seg(g, dt, B, cin, N, H, W, ks) for cin, ks in seg_spec]
    bias = torch.randn(B, N, generator=g) if per_sample else torch.randn(N, generator=g)
    res = torch.randn(B, N, H, W, generator=g) if with_res else None
    scale = 0.70710678
    plain, st_plain = run_conv_tc(dt, segs, B, H, W, N, bias, res, scale, want_stats=True)
    assert L.use_op_set_latency(1) == 0
    try:
        got, st = run_conv_tc(dt, segs, B, H, W, N, bias, res, scale, want_stats=True)
        got2, st2 = run_conv_tc(dt, segs, B, H, W, N, bias, res, scale, want_stats=True)
        one = run_conv_tc(dt, [(x[B - 1:], w, ks) for x, w, ks in segs], 1, H, W, N, bias[B - 1:] if per_sample else bias,
                          res[B - 1:] if res is not None else None, scale)
    finally:
        L.use_op_set_latency(0)
    ref = ref_conv(dt, segs, bias, res, scale)
    mx = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= OUT_TOL[dt] * mx, f"{name}: " + describe_mismatch(got, ref)
    # another association of the same fp32 products: the outputs agree to fp32 rounding (one output ulp in bf16)
    assert float((got - plain).abs().max()) <= (1e-5 if dt == F32 else 8e-3) * mx, describe_mismatch(got, plain)
    assert torch.equal(got, got2) and torch.equal(st, st2)          # deterministic
    assert torch.equal(got[B - 1:], one)                              # batch-invariant inside the mode
    assert torch.allclose(st, st_plain, rtol=1e-3, atol=1e-2 * mx)   # GroupNorm statistics of the output


@pytest.mark.parametrize("dt", [F32, BF16], ids=["tf32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 40, 20, 128, 128), (2, 24, 10, 256, 128), (1, 128, 160, 128, 128)])
def test_conv_tc_fused_groupnorm_stats(dt, shape):
    """The epilogue's per-channel sum / sum of squares of the STORED output (ragged tiles masked), bit-reproducible."""
    B, H, W, N, Cin = shape
    g = torch.Generator().manual_seed(17)
    segs = [make_seg(g, dt, B, Cin, N, H, W, 3)]
    bias = torch.randn(B, N, generator=g)
    got, stats = run_conv_tc(dt, segs, B, H, W, N, bias, None, 1.0, want_stats=True)
    got2, stats2 = run_conv_tc(dt, segs, B, H, W, N, bias, None, 1.0, want_stats=True)
    assert torch.equal(stats, stats2) and torch.equal(got, got2)  # integer atomics: bit-reproducible
    # the statistics describe the fp32 epilogue values (before the bf16 rounding of the store): compare with the fp64
    # reference convolution; bf16 rounding noise of `got` would be sqrt(n) * 2^-9 here
    ref = ref_conv(dt, segs, bias, None, 1.0).double()
    ref_sum, ref_sq = ref.sum(dim=(2, 3)), (ref ** 2).sum(dim=(2, 3))
    assert torch.allclose(stats[..., 0], ref_sum, rtol=1e-4, atol=2e-2), describe_mismatch(stats[..., 0], ref_sum)
    assert torch.allclose(stats[..., 1], ref_sq, rtol=1e-4, atol=2e-2), describe_mismatch(stats[..., 1], ref_sq)


@pytest.mark.parametrize("dt", [F32, BF16], ids=["tf32", "bf16"])
def test_conv_tc_weight_window(dt):
    """Two activation tensors against channel windows of ONE weight tensor (Conv_2 over cat[h, skip])."""
    L = _lib.lib()
    g = torch.Generator().manual_seed(5)
    B, H, W, N, C0, C1 = 1, 32, 16, 128, 128, 64
    x0 = to_operand(torch.randn(B, C0, H, W, generator=g), dt)
    x1 = to_operand(torch.randn(B, C1, H, W, generator=g), dt)
    w = to_operand(torch.randn(N, C0 + C1, 1, 1, generator=g) / np.sqrt(C0 + C1), dt)
    bias = torch.randn(N, generator=g)
    a0, a1, pw = act_tensor(x0, dt), act_tensor(x1, dt), pack_weight(L, w, dt)
    out = torch.empty(B, H, W, N, device="cuda", dtype=a0.dtype)
    bd = bias.cuda()
    rc = L.use_op_conv_tc(dt, 2, ptr_array([a0.data_ptr(), a1.data_ptr()]), int_array([C0, C1]), int_array([0, 0]),
                          int_array([C0, C1]), ptr_array([pw.data_ptr(), pw.data_ptr()]), int_array([C0 + C1, C0 + C1]),
                          int_array([0, C0]), int_array([1, 1]), B, H, W, N, bd.data_ptr(), 0, None, 1.0, out.data_ptr(),
                          None, stream())
    assert rc == 0, L.use_last_error()
    _sync()
    ref = Fnn.conv2d(torch.cat([x0, x1], 1).double(), w.double()).float() + bias[None, :, None, None]
    got = from_act(out)
    assert float((got - ref).abs().max()) <= OUT_TOL[dt] * float(ref.abs().max()), describe_mismatch(got, ref)


@pytest.mark.parametrize("dt", [F32, BF16], ids=["fp32", "bf16"])
def test_conv_ref_matches_torch(dt):
    L = _lib.lib()
    g = torch.Generator().manual_seed(1)
    B, H, W, Cin, Cout = 2, 9, 7, 8, 5
    x = to_operand(torch.randn(B, Cin, H, W, generator=g), dt)
    w = torch.randn(Cout, Cin, 3, 3, generator=g)
    bias = torch.randn(Cout, generator=g)
    a = act_tensor(x, dt)
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=a.dtype)
    wd, bd = w.cuda(), bias.cuda()
    assert L.use_op_conv_ref(dt, a.data_ptr(), wd.data_ptr(), bd.data_ptr(), 0, None, 1.0, out.data_ptr(), B, H, W, Cin,
                             Cout, 3, stream()) == 0
    _sync()
    ref = Fnn.conv2d(x.double(), w.double(), bias.double(), padding=1).float()
    assert float((from_act(out) - ref).abs().max()) <= OUT_TOL[dt] * float(ref.abs().max())


def stats_to_float(st_i64):
    """fixed point [.., 2] int64 -> float64 (sum, sum of squares)"""
    st = st_i64.cpu().to(torch.float64)
    return torch.stack([st[..., 0] / 2.0**28, st[..., 1] / 2.0**24], dim=-1)


def gn_stats_raw(L, dt, a, B, HW, Cc):
    st = torch.zeros(B, Cc, 2, dtype=torch.int64, device="cuda")
    assert L.use_op_gn_stats(dt, a.data_ptr(), st.data_ptr(), B, HW, Cc, stream()) == 0, L.use_last_error()
    torch.cuda.synchronize()
    return st


def gn_stats(L, dt, a, B, HW, Cc):
    return stats_to_float(gn_stats_raw(L, dt, a, B, HW, Cc))


def test_gn_stats_deterministic_and_batch_invariant():
    """No floating-point atomics: repeated launches and different batch sizes give bit-identical statistics."""
    L = _lib.lib()
    g = torch.Generator().manual_seed(0)
    B, H, W, Cc = 3, 96, 80, 128  # 7680 pixels -> 4 partial blocks per sample
    x = torch.randn(B, Cc, H, W, generator=g)
    for dt in (F32, BF16):
        a = act_tensor(x, dt)
        s1 = gn_stats_raw(L, dt, a, B, H * W, Cc)
        s2 = gn_stats_raw(L, dt, a, B, H * W, Cc)
        assert torch.equal(s1, s2)
        s3 = gn_stats_raw(L, dt, a[1:2].contiguous(), 1, H * W, Cc)
        assert torch.equal(s1[1:2], s3)
        ref = to_operand(x, dt).double().sum(dim=(2, 3)) if dt == BF16 else x.double().sum(dim=(2, 3))
        assert torch.allclose(stats_to_float(s1)[..., 0], ref, rtol=1e-5, atol=1e-2)


def _gn_ref(x, gamma, beta, silu):
    C = x.shape[1]
    y = Fnn.group_norm(x.double(), min(C // 4, 32), gamma.double(), beta.double(), eps=1e-6)
    return (Fnn.silu(y) if silu else y).float()


@pytest.mark.parametrize("dt", [F32, BF16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("C0,C1,fir", [(128, 0, 0), (64, 0, 0), (256, 128, 0), (128, 64, 0), (128, 0, 1), (128, 0, 2),
                                       (256, 0, 1)])
def test_groupnorm_silu_fir(dt, C0, C1, fir):
    """gn_stats + gn_apply vs torch group_norm -> SiLU -> FIR resample (layerspp.py:283-298)."""
    L = _lib.lib()
    g = torch.Generator().manual_seed(C0 + 7 * C1 + fir)
    B, H, W = 2, 12, 20
    x0 = to_operand(1.5 * torch.randn(B, C0, H, W, generator=g) + 0.3, dt) if dt == BF16 else 1.5 * torch.randn(B, C0, H, W, generator=g) + 0.3
    srcs = [x0]
    if C1:
        x1 = 0.7 * torch.randn(B, C1, H, W, generator=g) - 0.2
        srcs.append(to_operand(x1, dt) if dt == BF16 else x1)
    Ct = C0 + C1
    gamma, beta = 1 + 0.1 * torch.randn(Ct, generator=g), 0.1 * torch.randn(Ct, generator=g)
    acts = [act_tensor(s, dt) for s in srcs]
    stats = [gn_stats_raw(L, dt, a, B, H * W, s.shape[1]) for a, s in zip(acts, srcs)]
    _sync()
    # statistics themselves
    for st, s in zip(stats, srcs):
        ref_sum = s.double().sum(dim=(2, 3))
        ref_sq = (s.double() ** 2).sum(dim=(2, 3))
        stf = stats_to_float(st)
        assert torch.allclose(stf[..., 0], ref_sum, rtol=1e-5, atol=1e-3)
        assert torch.allclose(stf[..., 1], ref_sq, rtol=1e-5, atol=1e-3)
    Ho, Wo = (H // 2, W // 2) if fir == 1 else ((H * 2, W * 2) if fir == 2 else (H, W))
    out = torch.empty(B, Ho, Wo, Ct, device="cuda", dtype=acts[0].dtype)
    raw = torch.empty_like(out) if fir else None
    gd, bd = gamma.cuda(), beta.cuda()
    rc = L.use_op_gn_apply(dt, acts[0].data_ptr(), stats[0].data_ptr(), C0, acts[1].data_ptr() if C1 else None,
                           stats[1].data_ptr() if C1 else None, C1, gd.data_ptr(), bd.data_ptr(), 1e-6, fir, 1, 0,
                           out.data_ptr(), raw.data_ptr() if fir else None, B, H, W, stream())
    assert rc == 0, L.use_last_error()
    _sync()
    if fir and not C1:
        # the engine's form: scale / shift table precomputed once (use_op_gn_affine) -> bit-identical outputs
        afft = torch.empty(B, 2, C0, device="cuda", dtype=torch.float32)
        assert L.use_op_gn_affine(stats[0].data_ptr(), C0, None, 0, gd.data_ptr(), bd.data_ptr(), 1e-6, H * W, afft.data_ptr(), B,
                                  stream()) == 0
        out2, raw2 = torch.empty_like(out), torch.empty_like(out)
        rc = L.use_op_gn_apply_aff(dt, acts[0].data_ptr(), stats[0].data_ptr(), C0, None, None, 0, gd.data_ptr(), bd.data_ptr(),
                                   1e-6, fir, 1, 0, out2.data_ptr(), raw2.data_ptr(), B, H, W, afft.data_ptr(), stream())
        assert rc == 0, L.use_last_error()
        _sync()
        assert torch.equal(out, out2) and torch.equal(raw, raw2)
    xcat = torch.cat(srcs, 1)
    ref = _gn_ref(xcat, gamma, beta, True)
    ref_raw = xcat
    if fir == 1:
        ref, ref_raw = O.fir_downsample_2d(ref), O.fir_downsample_2d(ref_raw)
    elif fir == 2:
        ref, ref_raw = O.fir_upsample_2d(ref), O.fir_upsample_2d(ref_raw)
    tol = (1e-5 if dt == F32 else 6e-3) * float(ref.abs().max())
    got = from_act(out)
    assert float((got - ref).abs().max()) <= tol, describe_mismatch(got, ref)
    if fir:
        gr = from_act(raw)
        assert float((gr - ref_raw).abs().max()) <= (1e-5 if dt == F32 else 6e-3) * float(ref_raw.abs().max()), \
            describe_mismatch(gr, ref_raw)


FUSED_CASES = [
    # name, B, H, W, N, sources (channels of the concatenated raw tensors), extra 1x1 segments (channels)
    ("c128_n128", 2, 40, 20, 128, [128], []),
    ("c256_n128_ragged", 2, 20, 10, 128, [256], []),
    ("concat_256_128_n128", 1, 32, 24, 128, [256, 128], []),
    ("c256_n256", 2, 16, 24, 256, [256], []),
    ("concat_256_256_n256_small", 3, 8, 10, 256, [256, 256], []),
    ("conv1_with_skip_segments", 2, 32, 16, 128, [128], [256, 128]),
    ("many_tiles", 3, 96, 80, 128, [128], []),
]


@pytest.mark.parametrize("dt", [F32, BF16], ids=["tf32", "bf16"])
@pytest.mark.parametrize("case", FUSED_CASES, ids=[c[0] for c in FUSED_CASES])
def test_conv_tc_fused_groupnorm_operand(case, dt):
    """GroupNorm + SiLU applied inside the convolution's operand path (transform warps) == the separate gn_apply kernel
    followed by the same convolution, BIT FOR BIT (identical arithmetic, identical accumulation order), and both match
    torch group_norm -> silu -> conv2d within the operand-rounding tolerance."""
    name, B, H, W, N, srcC, extra = case
    L = _lib.lib()
    g = torch.Generator().manual_seed(len(name) * 31 + B)
    srcs = [to_operand((1.0 + 0.5 * i) * torch.randn(B, c, H, W, generator=g) + 0.2 * i, dt) for i, c in enumerate(srcC)]
    Ct = sum(srcC)
    gamma, beta = 1 + 0.1 * torch.randn(Ct, generator=g), 0.1 * torch.randn(Ct, generator=g)
    w3 = to_operand(torch.randn(N, Ct, 3, 3, generator=g) / np.sqrt(Ct * 9), dt)
    xs = [to_operand(torch.randn(B, c, H, W, generator=g), dt) for c in extra]
    w1 = to_operand(torch.randn(N, sum(extra), 1, 1, generator=g) / np.sqrt(max(1, sum(extra))), dt) if extra else None
    bias = torch.randn(B, N, generator=g)
    acts = [act_tensor(s, dt) for s in srcs]
    stats = [gn_stats_raw(L, dt, a, B, H * W, c) for a, c in zip(acts, srcC)]
    gd, bd = gamma.cuda(), beta.cuda()
    C0, C1 = srcC[0], (srcC[1] if len(srcC) > 1 else 0)
    # (1) unfused: gn_apply materialises the activated operand tensor, then a plain 3x3 segment over it
    a_act = torch.empty(B, H, W, Ct, device="cuda", dtype=acts[0].dtype)
    rc = L.use_op_gn_apply(dt, acts[0].data_ptr(), stats[0].data_ptr(), C0, acts[1].data_ptr() if C1 else None,
                           stats[1].data_ptr() if C1 else None, C1, gd.data_ptr(), bd.data_ptr(), 1e-6, 0, 1, 1,
                           a_act.data_ptr(), None, B, H, W, stream())
    assert rc == 0, L.use_last_error()
    pw3 = pack_weight(L, w3, dt)
    pw1 = pack_weight(L, w1, dt) if extra else None
    xacts = [act_tensor(x, dt) for x in xs]
    bias_d = bias.cuda()

    def run(fused):
        seg_act, ct, c0, cc, ws, cw, wc0, taps, aff, affc, affc0 = [], [], [], [], [], [], [], [], [], [], []
        afft = None
        if fused:
            afft = torch.empty(B, 2, Ct, device="cuda", dtype=torch.float32)
            rc = L.use_op_gn_affine(stats[0].data_ptr(), C0, stats[1].data_ptr() if C1 else None, C1, gd.data_ptr(),
                                    bd.data_ptr(), 1e-6, H * W, afft.data_ptr(), B, stream())
            assert rc == 0, L.use_last_error()
            off = 0
            for a, c in zip(acts, srcC):
                seg_act.append(a.data_ptr()); ct.append(c); c0.append(0); cc.append(c); ws.append(pw3.data_ptr()); cw.append(Ct)
                wc0.append(off); taps.append(9); aff.append(afft.data_ptr()); affc.append(Ct); affc0.append(off)
                off += c
        else:
            seg_act.append(a_act.data_ptr()); ct.append(Ct); c0.append(0); cc.append(Ct); ws.append(pw3.data_ptr()); cw.append(Ct)
            wc0.append(0); taps.append(9); aff.append(None); affc.append(0); affc0.append(0)
        if len(seg_act) + len(xacts) > 3:
            pytest.skip("more than 3 segments")
        off = 0
        for xa, c in zip(xacts, extra):
            seg_act.append(xa.data_ptr()); ct.append(c); c0.append(0); cc.append(c); ws.append(pw1.data_ptr()); cw.append(sum(extra))
            wc0.append(off); taps.append(1); aff.append(None); affc.append(0); affc0.append(0)
            off += c
        out = torch.empty(B, H, W, N, device="cuda", dtype=acts[0].dtype)
        st = torch.zeros(B, N, 2, dtype=torch.int64, device="cuda")
        rc = L.use_op_conv_tc_gn(dt, len(seg_act), ptr_array(seg_act), int_array(ct), int_array(c0), int_array(cc),
                                 ptr_array(ws), int_array(cw), int_array(wc0), int_array(taps), ptr_array(aff),
                                 int_array(affc), int_array(affc0), B, H, W, N, bias_d.data_ptr(), N, None, 0.70710678,
                                 out.data_ptr(), st.data_ptr(), stream())
        assert rc == 0, L.use_last_error()
        _sync()
        return out, st, afft

    out_u, st_u, _ = run(False)
    out_f, st_f, _ = run(True)
    assert torch.equal(out_u, out_f), describe_mismatch(from_act(out_f), from_act(out_u))
    assert torch.equal(st_u, st_f)
    ref_a = to_operand(_gn_ref(torch.cat(srcs, 1), gamma, beta, True), dt)
    ref = Fnn.conv2d(ref_a.double(), w3.double(), padding=1)
    if extra:
        ref = ref + Fnn.conv2d(torch.cat(xs, 1).double(), w1.double())
    ref = ((ref + bias.double()[:, :, None, None]) * 0.70710678).float()
    got = from_act(out_f)
    # the activated operand is re-rounded to bf16 / TF32: one operand ulp on a few inputs moves the sum slightly
    tol = (2e-3 if dt == F32 else 1.5e-2) * float(ref.abs().max())
    assert float((got - ref).abs().max()) <= tol, describe_mismatch(got, ref)


def test_gn_apply_operand_rounding_fp32():
    """as_operand=1 in fp32 mode stores TF32-representable values (low 13 mantissa bits clear)."""
    L = _lib.lib()
    B, H, W, Cc = 1, 4, 4, 64
    x = torch.randn(B, Cc, H, W)
    a = act_tensor(x, F32)
    st = gn_stats_raw(L, F32, a, B, H * W, Cc)
    out = torch.empty_like(a)
    gd, bd = torch.ones(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    assert L.use_op_gn_apply(F32, a.data_ptr(), st.data_ptr(), Cc, None, None, 0, gd.data_ptr(), bd.data_ptr(), 1e-6, 0, 1,
                             1, out.data_ptr(), None, B, H, W, stream()) == 0
    _sync()
    assert int((out.view(torch.int32) & 0x1FFF).abs().max()) == 0


@pytest.mark.parametrize("dt", [F32, BF16], ids=["fp32", "bf16"])
def test_conv_in4(dt):
    L = _lib.lib()
    g = torch.Generator().manual_seed(2)
    B, H, W, N = 2, 16, 24, 128
    x = torch.randn(B, 4, H, W, generator=g)
    w = torch.randn(N, 4, 3, 3, generator=g) / 6
    bias = torch.randn(N, generator=g)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    out = torch.empty(B, H, W, N, device="cuda", dtype=torch.bfloat16 if dt == BF16 else torch.float32)
    wd, bd = w.cuda(), bias.cuda()
    assert L.use_op_conv_in4(dt, xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(), B, H, W, N, stream()) == 0
    _sync()
    ref = Fnn.conv2d(x.double(), w.double(), bias.double(), padding=1).float()
    got = from_act(out)
    assert float((got - ref).abs().max()) <= OUT_TOL[dt] * float(ref.abs().max()), describe_mismatch(got, ref)


@pytest.mark.parametrize("dt", [F32, BF16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("with_prev", [False, True])
def test_conv_out4(dt, with_prev):
    """pyramid conv3x3 C->4 (+ FIR-up of the previous pyramid): ncsnpp.py:440-461."""
    L = _lib.lib()
    g = torch.Generator().manual_seed(3)
    B, H, W, Cc = 2, 16, 20, 128
    a = to_operand(torch.randn(B, Cc, H, W, generator=g), dt) if dt == BF16 else torch.randn(B, Cc, H, W, generator=g)
    w = torch.randn(4, Cc, 3, 3, generator=g) / np.sqrt(9 * Cc)
    bias = torch.randn(4, generator=g)
    prev = torch.randn(B, 4, H // 2, W // 2, generator=g)
    ad = act_tensor(a, dt)
    wd, bd = w.cuda(), bias.cuda()
    pd = prev.permute(0, 2, 3, 1).contiguous().cuda()
    out = torch.empty(B, H, W, 4, device="cuda")
    assert L.use_op_conv_out4(dt, ad.data_ptr(), wd.data_ptr(), bd.data_ptr(), pd.data_ptr() if with_prev else None,
                              out.data_ptr(), B, H, W, Cc, stream()) == 0
    _sync()
    ref = Fnn.conv2d(a.double(), w.double(), bias.double(), padding=1).float()
    if with_prev:
        ref = ref + O.fir_upsample_2d(prev)
    got = out.permute(0, 3, 1, 2).cpu()
    assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max()), describe_mismatch(got, ref)


HEAD_CASES = [
    # B, H, W, C, pc, with_prev
    (2, 16, 20, 128, 4, False),
    (2, 16, 20, 128, 4, True),
    (1, 8, 1, 256, 4, True),       # bottom of a short-clip pyramid: a single column
    (3, 28, 42, 256, 4, True),     # exact multiples of the 14 x 14 tile
    (2, 64, 80, 128, 4, True),     # many tiles, persistent loop + accumulator double buffering
    (2, 30, 18, 128, 2, True),     # 2-channel head (discriminative generator)
]


@pytest.mark.parametrize("dt", [F32, BF16], ids=["tf32", "bf16"])
@pytest.mark.parametrize("case", HEAD_CASES, ids=[f"B{c[0]}_{c[1]}x{c[2]}_C{c[3]}_pc{c[4]}_{'prev' if c[5] else 'noprev'}" for c in HEAD_CASES])
def test_head_tc(case, dt):
    """pyramid head on the tensor cores with the nine taps folded into N (head_tc.cuh) vs torch conv2d (+ FIR-up)."""
    B, H, W, Cc, pc, with_prev = case
    L = _lib.lib()
    g = torch.Generator().manual_seed(B * 1000 + H + W + Cc)
    a = to_operand(torch.randn(B, Cc, H, W, generator=g), dt)
    w = torch.randn(pc, Cc, 3, 3, generator=g) / np.sqrt(9 * Cc)
    bias = torch.randn(pc, generator=g)
    prev = torch.randn(B, pc, H // 2, W // 2, generator=g) if with_prev and H % 2 == 0 and W % 2 == 0 else None
    ad = act_tensor(a, dt)
    wh = w.contiguous()
    bd = bias.cuda()
    pd = prev.permute(0, 2, 3, 1).contiguous().cuda() if prev is not None else None
    out = torch.full((B, H, W, pc), float("nan"), device="cuda")
    scratch = torch.empty(48 * Cc * 4, dtype=torch.uint8, device="cuda")
    rc = L.use_op_head_tc(dt, ad.data_ptr(), wh.data_ptr(), bd.data_ptr(), pd.data_ptr() if pd is not None else None,
                          out.data_ptr(), B, H, W, Cc, pc, scratch.data_ptr(), stream())
    assert rc == 0, L.use_last_error()
    _sync()
    ref = Fnn.conv2d(a.double(), to_operand(w, dt).double(), bias.double(), padding=1).float()
    if prev is not None:
        ref = ref + O.fir_upsample_2d(prev)
    got = out.permute(0, 3, 1, 2).cpu()
    assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max()), describe_mismatch(got, ref)


@pytest.mark.parametrize("dt", [F32, BF16], ids=["tf32", "bf16"])
@pytest.mark.parametrize("case", HEAD_CASES, ids=[f"B{c[0]}_{c[1]}x{c[2]}_C{c[3]}_pc{c[4]}_{'prev' if c[5] else 'noprev'}" for c in HEAD_CASES])
def test_head_tc_fused_groupnorm_operand(case, dt):
    """Pyramid head with GroupNorm + SiLU applied inside its operand path (transform warps, head_tc.cuh FUSE) == gn_apply
    followed by the plain head, BIT FOR BIT, and both match torch group_norm -> silu -> conv2d (+ FIR-up) within the
    operand-rounding tolerance (ncsnpp.py:440-461)."""
    B, H, W, Cc, pc, with_prev = case
    L = _lib.lib()
    g = torch.Generator().manual_seed(B * 977 + H + W + Cc + pc)
    x = to_operand(1.5 * torch.randn(B, Cc, H, W, generator=g) + 0.3, dt)
    gamma, beta = 1 + 0.1 * torch.randn(Cc, generator=g), 0.1 * torch.randn(Cc, generator=g)
    w = torch.randn(pc, Cc, 3, 3, generator=g) / np.sqrt(9 * Cc)
    bias = torch.randn(pc, generator=g)
    prev = torch.randn(B, pc, H // 2, W // 2, generator=g) if with_prev and H % 2 == 0 and W % 2 == 0 else None
    xd = act_tensor(x, dt)
    st = gn_stats_raw(L, dt, xd, B, H * W, Cc)
    gd, btd, bd = gamma.cuda(), beta.cuda(), bias.cuda()
    wh = w.contiguous()
    pd = prev.permute(0, 2, 3, 1).contiguous().cuda() if prev is not None else None
    scratch = torch.empty(48 * Cc * 4, dtype=torch.uint8, device="cuda")
    # (1) unfused: gn_apply materialises the activated operand, then the plain head
    a_act = torch.empty_like(xd)
    rc = L.use_op_gn_apply(dt, xd.data_ptr(), st.data_ptr(), Cc, None, None, 0, gd.data_ptr(), btd.data_ptr(), 1e-6, 0, 1, 1,
                           a_act.data_ptr(), None, B, H, W, stream())
    assert rc == 0, L.use_last_error()
    out_u = torch.full((B, H, W, pc), float("nan"), device="cuda")
    rc = L.use_op_head_tc(dt, a_act.data_ptr(), wh.data_ptr(), bd.data_ptr(), pd.data_ptr() if pd is not None else None,
                          out_u.data_ptr(), B, H, W, Cc, pc, scratch.data_ptr(), stream())
    assert rc == 0, L.use_last_error()
    # (2) fused: the head reads the raw tensor + the scale / shift table
    afft = torch.empty(B, 2, Cc, device="cuda", dtype=torch.float32)
    rc = L.use_op_gn_affine(st.data_ptr(), Cc, None, 0, gd.data_ptr(), btd.data_ptr(), 1e-6, H * W, afft.data_ptr(), B, stream())
    assert rc == 0, L.use_last_error()
    out_f = torch.full((B, H, W, pc), float("nan"), device="cuda")
    rc = L.use_op_head_tc_gn(dt, xd.data_ptr(), afft.data_ptr(), wh.data_ptr(), bd.data_ptr(),
                             pd.data_ptr() if pd is not None else None, out_f.data_ptr(), B, H, W, Cc, pc, scratch.data_ptr(),
                             stream())
    assert rc == 0, L.use_last_error()
    _sync()
    assert torch.equal(out_u, out_f), describe_mismatch(out_f.cpu(), out_u.cpu())
    ref_a = to_operand(_gn_ref(x, gamma, beta, True), dt)
    ref = Fnn.conv2d(ref_a.double(), to_operand(w, dt).double(), bias.double(), padding=1).float()
    if prev is not None:
        ref = ref + O.fir_upsample_2d(prev)
    got = out_f.permute(0, 3, 1, 2).cpu()
    tol = (2e-3 if dt == F32 else 1.5e-2) * float(ref.abs().max())
    assert float((got - ref).abs().max()) <= tol, describe_mismatch(got, ref)


@pytest.mark.parametrize("dt", [F32, BF16], ids=["fp32", "bf16"])
def test_combine_and_fir4(dt):
    L = _lib.lib()
    g = torch.Generator().manual_seed(4)
    B, H, W, Cc = 2, 16, 24, 128
    pyr = torch.randn(B, 4, H, W, generator=g)
    pd = pyr.permute(0, 2, 3, 1).contiguous().cuda()
    pdown = torch.empty(B, H // 2, W // 2, 4, device="cuda")
    assert L.use_op_fir4_down(pd.data_ptr(), pdown.data_ptr(), B, H, W, 4, stream()) == 0
    _sync()
    ref_down = O.fir_downsample_2d(pyr)
    assert float((pdown.permute(0, 3, 1, 2).cpu() - ref_down).abs().max()) < 1e-5
    h = to_operand(torch.randn(B, Cc, H // 2, W // 2, generator=g), dt) if dt == BF16 else torch.randn(B, Cc, H // 2, W // 2, generator=g)
    w = torch.randn(Cc, 4, 1, 1, generator=g) / 2
    bias = torch.randn(Cc, generator=g)
    hd = act_tensor(h, dt)
    wd, bd = w.cuda(), bias.cuda()
    assert L.use_op_combine(dt, hd.data_ptr(), pdown.data_ptr(), wd.data_ptr(), bd.data_ptr(), hd.data_ptr(), B,
                            (H // 2) * (W // 2), Cc, 4, stream()) == 0
    _sync()
    ref = Fnn.conv2d(ref_down.double(), w.double(), bias.double()).float() + h
    got = from_act(hd)
    assert float((got - ref).abs().max()) <= OUT_TOL[dt] * float(ref.abs().max()), describe_mismatch(got, ref)


@pytest.mark.parametrize("dt", [F32, BF16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("Cc,pc,HW", [(128, 4, (48, 40)), (256, 4, (30, 36)), (256, 2, (64, 20))])
def test_combine_with_statistics(dt, Cc, pc, HW):
    """Combine + fused GroupNorm statistics of its output: values vs torch, statistics vs the fp64 sums, bit-reproducible
    and batch-invariant (fixed block partition, integer atomics)."""
    L = _lib.lib()
    g = torch.Generator().manual_seed(Cc + pc)
    B, (H, W) = 3, HW
    pyr = torch.randn(B, pc, H, W, generator=g)
    h = to_operand(torch.randn(B, Cc, H, W, generator=g), dt)
    w = torch.randn(Cc, pc, 1, 1, generator=g) / 2
    bias = torch.randn(Cc, generator=g)
    pd = pyr.permute(0, 2, 3, 1).contiguous().cuda()
    hd = act_tensor(h, dt)
    wd, bd = w.cuda(), bias.cuda()

    def run(hh, pp, nb):
        out = torch.empty_like(hh)
        st = torch.zeros(nb, Cc, 2, dtype=torch.int64, device="cuda")
        rc = L.use_op_combine_stats(dt, hh.data_ptr(), pp.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(),
                                    st.data_ptr(), nb, H * W, Cc, pc, stream())
        assert rc == 0, L.use_last_error()
        _sync()
        return out, st

    out, st = run(hd, pd, B)
    out2, st2 = run(hd, pd, B)
    assert torch.equal(out, out2) and torch.equal(st, st2)
    out1, st1 = run(hd[1:2].contiguous(), pd[1:2].contiguous(), 1)
    assert torch.equal(out[1:2], out1) and torch.equal(st[1:2], st1)
    ref = Fnn.conv2d(pyr.double(), w.double(), bias.double()).float() + h
    got = from_act(out)
    assert float((got - ref).abs().max()) <= OUT_TOL[dt] * float(ref.abs().max()), describe_mismatch(got, ref)
    stf = stats_to_float(st)
    assert torch.allclose(stf[..., 0], ref.double().sum(dim=(2, 3)), rtol=1e-4, atol=2e-2)
    assert torch.allclose(stf[..., 1], (ref.double() ** 2).sum(dim=(2, 3)), rtol=1e-4, atol=2e-2)


@pytest.mark.parametrize("up,down,pad", [(2, 1, (2, 1)), (1, 2, (1, 1)), (1, 1, (0, 0)), (2, 2, (1, 0))])
def test_upfirdn2d_abi(up, down, pad):
    """The reference's native op seam (op/upfirdn2d.cpp:12-23) against its CPU semantics."""
    L = _lib.lib()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(3, 2, 9, 11, generator=g)  # [N, C, H, W] -> planes [N*C, H, W, 1]
    k = torch.tensor(O._setup_kernel((1, 3, 3, 1)) * (up**2))
    ref = O._upfirdn2d(x, k, up, down, pad[0], pad[1])
    oh, ow = ref.shape[-2:]
    xd, kd = x.reshape(-1, 9, 11, 1).contiguous().cuda(), k.cuda()
    out = torch.empty(6, oh, ow, 1, device="cuda")
    assert L.use_upfirdn2d_f32(xd.data_ptr(), out.data_ptr(), 6, 9, 11, 1, kd.data_ptr(), 4, 4, up, up, down, down, pad[0],
                               pad[1], pad[0], pad[1], stream()) == 0
    _sync()
    assert float((out.reshape(3, 2, oh, ow).cpu() - ref).abs().max()) < 1e-5


def test_philox_complex_normal_statistics():
    """In-kernel noise: complex standard normal (Re, Im ~ N(0, 1/2)), independent across step / clip streams."""
    L = _lib.lib()
    B, per = 4, 1 << 18
    z = torch.empty(B, per, dtype=torch.complex64, device="cuda")
    assert L.use_op_philox(z.data_ptr(), 1234, 3, 10, B, per, stream()) == 0
    z2 = torch.empty_like(z)
    assert L.use_op_philox(z2.data_ptr(), 1234, 4, 10, B, per, stream()) == 0
    z3 = torch.empty(2, per, dtype=torch.complex64, device="cuda")
    assert L.use_op_philox(z3.data_ptr(), 1234, 3, 12, 2, per, stream()) == 0  # clips 12, 13 = rows 2, 3 of z
    _sync()
    zr = torch.view_as_real(z).double().cpu()
    n = zr.numel()
    assert abs(float(zr.mean())) < 5 / np.sqrt(n)
    assert abs(float(zr.var()) - 0.5) < 5 * 0.5 * np.sqrt(2 / n)
    assert abs(float((zr**4).mean()) - 3 * 0.25) < 0.01                       # kurtosis of N(0, 1/2)
    assert abs(float((zr[..., 0] * zr[..., 1]).mean())) < 5 * 0.5 / np.sqrt(n / 2)  # Re / Im uncorrelated
    assert abs(float((zr * torch.view_as_real(z2).double().cpu()).mean())) < 5 * 0.5 / np.sqrt(n)
    assert torch.equal(z[2:], z3)  # shard invariance: stream depends on the global clip index only
