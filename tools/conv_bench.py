#!/usr/bin/env python
"""Kernel-tuning tool (not product code): times single tcgen05 convolution launches of the shapes NCSNppLarge runs,
through the C ABI (use_op_conv_tc_gn with the USE_B200_CONV_TIME hook), fused GroupNorm operand on / off.

  python tools/conv_bench.py [--batch 16] [--reps 10] [--dtype fp32|bf16|both]
Prints one line per case: ms per launch and TFLOP/s.  USE_B200_CONV_DBG selects experiment switches of the kernel.
"""
import argparse
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [
    # name, H, W, N, fused sources, TMA 3x3 sources, 1x1 sources
    ("L0 conv0 128->128", 512, 640, 128, [128], [], []),
    ("L0 conv1 128->128 +res", 512, 640, 128, [128], [], []),
    ("L0up conv0 cat(128,128)->128", 512, 640, 128, [128, 128], [], []),
    ("L0up conv1 128->128 + 1x1(128,128)", 512, 640, 128, [128], [], [128, 128]),
    ("L0up conv0 cat(256,128)->128", 512, 640, 128, [256, 128], [], []),
    ("L2 conv0 256->256", 128, 160, 256, [256], [], []),
    ("L2up conv0 cat(256,256)->256", 128, 160, 256, [256, 256], [], []),
    ("L2up conv1 256->256 + 1x1(256,256)", 128, 160, 256, [256], [], [256, 256]),
]


SMALL_CASES = [
    ("L6 conv 256->256 (8x10)", 8, 10, 256, [256], [], []),
    ("L6 conv 64->256 (8x10)", 8, 10, 256, [64], [], []),
    ("L6 conv cat(256,256)->256 (8x10)", 8, 10, 256, [256, 256], [], []),
    ("L5 conv 256->256 (16x20)", 16, 20, 256, [256], [], []),
    ("L4 conv 256->256 (32x40)", 32, 40, 256, [256], [], []),
    ("L3 conv 256->256 (64x80)", 64, 80, 256, [256], [], []),
    ("L3 conv 64->256 (64x80)", 64, 80, 256, [64], [], []),
    ("L3 conv cat(256,256)->256 (64x80)", 64, 80, 256, [256, 256], [], []),
    ("L2 conv 256->256 (128x160)", 128, 160, 256, [256], [], []),
]


def main():
    global CASES
    if os.environ.get("CONV_BENCH_SMALL"):
        CASES = SMALL_CASES
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--dtype", default="both")
    args = ap.parse_args()
    os.environ["USE_B200_CONV_TIME"] = str(args.reps)
    # the timing hook prints from C to stderr: point fd 2 at a file, run every case, parse the lines in order
    log = os.path.join(ROOT, "gpurun_out", "conv_bench_stderr.txt")
    os.makedirs(os.path.dirname(log), exist_ok=True)
    fd = os.open(log, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    os.dup2(fd, 2)

    import torch
    from use_b200 import _lib
    from util import BF16, F32, int_array, ptr_array, stream

    L = _lib.lib()

    def run_case(dtn, ci, fused):
        dt = BF16 if dtn == "bf16" else F32
        name, H, W, N, fs, ts, os_ = CASES[ci]
        B = args.batch
        tdt = torch.bfloat16 if dt == BF16 else torch.float32
        es = 2 if dt == BF16 else 4
        g = torch.Generator(device="cuda").manual_seed(0)

        def rnd(*shape):
            return torch.randn(*shape, device="cuda", generator=g, dtype=torch.float32).to(tdt)

        seg_act, ct, c0, cc, ws, cw, wc0, taps, aff, affc, affc0, keep = [], [], [], [], [], [], [], [], [], [], [], []
        C3 = sum(fs) + sum(ts)
        w3 = torch.zeros(9 * N * C3 * es, dtype=torch.uint8, device="cuda")
        off = 0
        Cf = sum(fs)
        afft = torch.ones(B, 2, max(Cf, 1), device="cuda", dtype=torch.float32)
        for c in fs:
            a = rnd(B, H, W, c)
            keep.append(a)
            seg_act.append(a.data_ptr()); ct.append(c); c0.append(0); cc.append(c); ws.append(w3.data_ptr()); cw.append(C3)
            wc0.append(off); taps.append(9)
            aff.append(afft.data_ptr() if fused else None); affc.append(Cf if fused else 0); affc0.append(off if fused else 0)
            off += c
        if not fused and len(fs) > 1:
            # the unfused path convolves ONE materialised concatenated tensor
            seg_act, ct, c0, cc, ws, cw, wc0, taps, aff, affc, affc0 = [], [], [], [], [], [], [], [], [], [], []
            a = rnd(B, H, W, Cf)
            keep.append(a)
            seg_act.append(a.data_ptr()); ct.append(Cf); c0.append(0); cc.append(Cf); ws.append(w3.data_ptr()); cw.append(C3)
            wc0.append(0); taps.append(9); aff.append(None); affc.append(0); affc0.append(0)
        C1 = sum(os_)
        w1 = torch.zeros(max(1, N * C1 * es), dtype=torch.uint8, device="cuda")
        off = 0
        for c in os_:
            a = rnd(B, H, W, c)
            keep.append(a)
            seg_act.append(a.data_ptr()); ct.append(c); c0.append(0); cc.append(c); ws.append(w1.data_ptr()); cw.append(C1)
            wc0.append(off); taps.append(1); aff.append(None); affc.append(0); affc0.append(0)
            off += c
        out = torch.empty(B, H, W, N, device="cuda", dtype=tdt)
        res = rnd(B, H, W, N) if "+res" in name else None
        st = torch.zeros(B, N, 2, dtype=torch.int64, device="cuda")
        bias = torch.zeros(B, N, device="cuda")
        torch.cuda.synchronize()
        rc = L.use_op_conv_tc_gn(dt, len(seg_act), ptr_array(seg_act), int_array(ct), int_array(c0), int_array(cc), ptr_array(ws),
                                 int_array(cw), int_array(wc0), int_array(taps), ptr_array(aff), int_array(affc), int_array(affc0),
                                 B, H, W, N, bias.data_ptr(), N, res.data_ptr() if res is not None else None, 0.7071,
                                 out.data_ptr(), st.data_ptr(), stream())
        assert rc == 0, L.use_last_error()
        torch.cuda.synchronize()

    def run_head(dtn, H, W, Cc, fused=False):
        dt = BF16 if dtn == "bf16" else F32
        B = args.batch
        tdt = torch.bfloat16 if dt == BF16 else torch.float32
        a = torch.randn(B, H, W, Cc, device="cuda").to(tdt)
        w = torch.randn(4, Cc, 3, 3)
        bias = torch.zeros(4, device="cuda")
        prev = torch.randn(B, H // 2, W // 2, 4, device="cuda")
        out = torch.empty(B, H, W, 4, device="cuda")
        scratch = torch.empty(48 * Cc * 4, dtype=torch.uint8, device="cuda")
        if fused:
            afft = torch.ones(B, 2, Cc, device="cuda", dtype=torch.float32)
            rc = L.use_op_head_tc_gn(dt, a.data_ptr(), afft.data_ptr(), w.data_ptr(), bias.data_ptr(), prev.data_ptr(), out.data_ptr(),
                                     B, H, W, Cc, 4, scratch.data_ptr(), stream())
        else:
            rc = L.use_op_head_tc(dt, a.data_ptr(), w.data_ptr(), bias.data_ptr(), prev.data_ptr(), out.data_ptr(), B, H, W, Cc, 4,
                                  scratch.data_ptr(), stream())
        assert rc == 0, L.use_last_error()
        torch.cuda.synchronize()

    def run_fir(dtn, H, W, Cc, fir):
        dt = BF16 if dtn == "bf16" else F32
        B = args.batch
        tdt = torch.bfloat16 if dt == BF16 else torch.float32
        x = torch.randn(B, H, W, Cc, device="cuda").to(tdt)
        st = torch.zeros(B, Cc, 2, dtype=torch.int64, device="cuda")
        assert L.use_op_gn_stats(dt, x.data_ptr(), st.data_ptr(), B, H * W, Cc, stream()) == 0
        Ho, Wo = (H // 2, W // 2) if fir == 1 else ((H * 2, W * 2) if fir == 2 else (H, W))
        out = torch.empty(B, Ho, Wo, Cc, device="cuda", dtype=tdt)
        raw = torch.empty_like(out) if fir else None
        g, bt = torch.ones(Cc, device="cuda"), torch.zeros(Cc, device="cuda")

        afft = torch.empty(B, 2, Cc, device="cuda", dtype=torch.float32)
        assert L.use_op_gn_affine(st.data_ptr(), Cc, None, 0, g.data_ptr(), bt.data_ptr(), 1e-6, H * W, afft.data_ptr(), B, stream()) == 0

        def call():
            rc = L.use_op_gn_apply_aff(dt, x.data_ptr(), st.data_ptr(), Cc, None, None, 0, g.data_ptr(), bt.data_ptr(), 1e-6, fir, 1, 1,
                                       out.data_ptr(), raw.data_ptr() if fir else None, B, H, W, afft.data_ptr() if fir else None,
                                       stream())
            assert rc == 0, L.use_last_error()

        call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        es = 2 if dt == BF16 else 4
        nbytes = B * H * W * Cc * es + (2 if fir else 1) * B * Ho * Wo * Cc * es
        print(f"{dtn} gn_apply fir={fir} {H}x{W} C{Cc}: {ms:7.3f} ms  {nbytes / ms / 1e6:7.1f} GB/s (algorithmic)", flush=True)

    dts = ["fp32", "bf16"] if args.dtype == "both" else [args.dtype]
    if os.environ.get("CONV_BENCH_ONLY_HEAD"):
        cases = [(dtn, H, W, Cc, fused) for dtn in dts for (H, W, Cc) in ((512, 640, 128), (256, 320, 128), (64, 80, 256))
                 for fused in (False, True)]
        for c in cases:
            run_head(*c)
        ms = [float(l.split("ms_per_launch=")[1].split()[0]) for l in open(log) if "USE_B200_CONV_TIME" in l]
        for (dtn, H, W, Cc, fused), t in zip(cases, ms):
            gb = args.batch * H * W * Cc * (2 if dtn == "bf16" else 4) / 1e9
            print(f"{dtn} head {H}x{W} C{Cc} {'fused' if fused else 'plain'}: {t:7.3f} ms  {gb / t * 1e3:7.1f} GB/s (operand read)", flush=True)
        return
    if os.environ.get("CONV_BENCH_ONLY_FIR"):
        for dtn in dts:
            for (H, W, Cc, fir) in ((512, 640, 128, 1), (256, 320, 128, 1), (128, 160, 256, 1), (256, 320, 128, 2), (128, 160, 128, 2),
                                    (512, 640, 128, 0)):
                run_fir(dtn, H, W, Cc, fir)
        return
    for dtn in dts:
        for ci in range(len(CASES)):
            for fused in (0, 1):
                run_case(dtn, ci, fused)
    if not os.environ.get("CONV_BENCH_SMALL"):
        for dtn in dts:
            run_head(dtn, 512, 640, 128)
            run_head(dtn, 256, 320, 128)
    report(args)


def report(args):
    log = os.path.join(ROOT, "gpurun_out", "conv_bench_stderr.txt")
    ms = [float(l.split("ms_per_launch=")[1].split()[0]) for l in open(log) if "USE_B200_CONV_TIME" in l]
    k = 0
    for dtn in (["fp32", "bf16"] if args.dtype == "both" else [args.dtype]):
        for case in CASES:
            name, H, W, N, fs, ts, os_ = case
            flops = 2.0 * args.batch * H * W * N * (9 * (sum(fs) + sum(ts)) + sum(os_))
            line = f"{dtn} {name:40s}"
            for fused in (0, 1):
                if k < len(ms):
                    line += f" | {'fused' if fused else 'plain'} {ms[k]:7.3f} ms {flops / ms[k] / 1e9:7.1f} TF/s"
                k += 1
            print(line, flush=True)
    for dtn in ([] if os.environ.get("CONV_BENCH_SMALL") else (["fp32", "bf16"] if args.dtype == "both" else [args.dtype])):
        for (H, W, Cc) in ((512, 640, 128), (256, 320, 128)):
            if k < len(ms):
                gb = args.batch * H * W * Cc * (2 if dtn == "bf16" else 4) / 1e9
                print(f"{dtn} head {H}x{W} C{Cc}: {ms[k]:7.3f} ms  {gb / ms[k] * 1e3:7.1f} GB/s (operand read)", flush=True)
            k += 1


if __name__ == "__main__":
    main()
