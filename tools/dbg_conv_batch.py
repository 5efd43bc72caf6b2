import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from util import BF16, F32
import test_gpu_kernels as K
for dt in (BF16, F32):
    for (H, W, N, Cin) in [(8, 10, 256, 256), (16, 20, 256, 512), (32, 40, 256, 256), (64, 80, 256, 256), (16, 24, 128, 128)]:
        g = torch.Generator().manual_seed(3)
        B = 3
        segs = [K.make_seg(g, dt, B, Cin, N, H, W, 3)]
        bias = torch.randn(B, N, generator=g)
        res = torch.randn(B, N, H, W, generator=g)
        full, st = K.run_conv_tc(dt, segs, B, H, W, N, bias, res, 0.7, want_stats=True)
        for b in range(B):
            s1 = [(x[b:b+1], w, ks) for x, w, ks in segs]
            one, st1 = K.run_conv_tc(dt, s1, 1, H, W, N, bias[b:b+1], res[b:b+1], 0.7, want_stats=True)
            print(dt, (H, W, N, Cin), "b", b, "out equal", torch.equal(one[0], full[b]), "stats equal", torch.equal(st1[0], st[b]),
                  float((one[0]-full[b]).abs().max()))
