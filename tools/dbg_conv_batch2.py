"""fused-GN conv: batch of 3 vs each sample alone, at the 64x80 level (unit counts 480 vs 160)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from util import BF16, F32, act_tensor, to_operand, pack_weight, ptr_array, int_array, stream
from use_b200 import _lib
import test_gpu_kernels as K
L = _lib.lib()

def conv(dt, srcs, srcC, gamma, beta, w3, xs, extra, w1, bias, B, H, W, N, fused):
    acts = [act_tensor(s, dt) for s in srcs]
    stats = [K.gn_stats_raw(L, dt, a, B, H * W, c) for a, c in zip(acts, srcC)]
    gd, bd = gamma.cuda(), beta.cuda()
    Ct = sum(srcC); C0, C1 = srcC[0], (srcC[1] if len(srcC) > 1 else 0)
    pw3 = pack_weight(L, w3, dt); pw1 = pack_weight(L, w1, dt) if extra else None
    xacts = [act_tensor(x, dt) for x in xs]
    bias_d = bias.cuda()
    seg_act, ct, c0, cc, ws, cw, wc0, taps, aff, affc, affc0 = [], [], [], [], [], [], [], [], [], [], []
    afft = torch.empty(B, 2, Ct, device="cuda", dtype=torch.float32)
    assert L.use_op_gn_affine(stats[0].data_ptr(), C0, stats[1].data_ptr() if C1 else None, C1, gd.data_ptr(), bd.data_ptr(), 1e-6, H * W, afft.data_ptr(), B, stream()) == 0
    if fused:
        off = 0
        for a, c in zip(acts, srcC):
            seg_act.append(a.data_ptr()); ct.append(c); c0.append(0); cc.append(c); ws.append(pw3.data_ptr()); cw.append(Ct)
            wc0.append(off); taps.append(9); aff.append(afft.data_ptr()); affc.append(Ct); affc0.append(off); off += c
    else:
        a_act = torch.empty(B, H, W, Ct, device="cuda", dtype=acts[0].dtype)
        assert L.use_op_gn_apply(dt, acts[0].data_ptr(), stats[0].data_ptr(), C0, acts[1].data_ptr() if C1 else None, stats[1].data_ptr() if C1 else None, C1, gd.data_ptr(), bd.data_ptr(), 1e-6, 0, 1, 1, a_act.data_ptr(), None, B, H, W, stream()) == 0
        seg_act.append(a_act.data_ptr()); ct.append(Ct); c0.append(0); cc.append(Ct); ws.append(pw3.data_ptr()); cw.append(Ct)
        wc0.append(0); taps.append(9); aff.append(None); affc.append(0); affc0.append(0)
    off = 0
    for xa, c in zip(xacts, extra):
        seg_act.append(xa.data_ptr()); ct.append(c); c0.append(0); cc.append(c); ws.append(pw1.data_ptr()); cw.append(sum(extra))
        wc0.append(off); taps.append(1); aff.append(None); affc.append(0); affc0.append(0); off += c
    out = torch.empty(B, H, W, N, device="cuda", dtype=acts[0].dtype)
    st = torch.zeros(B, N, 2, dtype=torch.int64, device="cuda")
    rc = L.use_op_conv_tc_gn(dt, len(seg_act), ptr_array(seg_act), int_array(ct), int_array(c0), int_array(cc), ptr_array(ws), int_array(cw), int_array(wc0), int_array(taps), ptr_array(aff), int_array(affc), int_array(affc0), B, H, W, N, bias_d.data_ptr(), N, None, 0.70710678, out.data_ptr(), st.data_ptr(), stream())
    assert rc == 0, L.use_last_error()
    torch.cuda.synchronize()
    return out, st

for dt in (BF16, F32):
  for (H, W, N, srcC, extra) in [(64, 80, 256, [256], []), (64, 80, 256, [256, 256], []), (64, 80, 256, [256], [256, 256]), (64, 80, 256, [256], [256])]:
    for fused in (True, False):
        if fused and len(srcC) + len(extra) > 3: continue
        g = torch.Generator().manual_seed(7)
        B = 3
        srcs = [to_operand(torch.randn(B, c, H, W, generator=g), dt) for c in srcC]
        Ct = sum(srcC)
        gamma, beta = 1 + 0.1 * torch.randn(Ct, generator=g), 0.1 * torch.randn(Ct, generator=g)
        w3 = to_operand(torch.randn(N, Ct, 3, 3, generator=g) / np.sqrt(Ct * 9), dt)
        xs = [to_operand(torch.randn(B, c, H, W, generator=g), dt) for c in extra]
        w1 = to_operand(torch.randn(N, sum(extra), 1, 1, generator=g) / np.sqrt(max(1, sum(extra))), dt) if extra else None
        bias = torch.randn(B, N, generator=g)
        full, st = conv(dt, srcs, srcC, gamma, beta, w3, xs, extra, w1, bias, B, H, W, N, fused)
        eq = []
        for b in range(B):
            one, st1 = conv(dt, [s[b:b+1] for s in srcs], srcC, gamma, beta, w3, [x[b:b+1] for x in xs], extra, w1, bias[b:b+1], 1, H, W, N, fused)
            eq.append((bool(torch.equal(one[0], full[b])), bool(torch.equal(st1[0], st[b]))))
        print("dt", dt, (H, W, N, srcC, extra), "fused", fused, eq)
