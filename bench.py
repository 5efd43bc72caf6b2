#!/usr/bin/env python
"""bench.py -- throughput of the SGMSE reverse-SDE sampling path (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic 4 s @ 24 kHz clips:
STFT + compression -> prior draw -> N x (NCSN++ Large evaluation + fused reverse-diffusion update) -> iSTFT.
Default workload = BASELINE.json configs[1]: batch 32 per GPU, N = 30, fp32 storage with TF32 tensor-core
convolutions (PyTorch's own GPU default for the reference's fp32 model).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|3|4] [--dtype fp32|bf16] [--batch B] [--N 30]
  python bench.py --impl reference ...     # the reference algorithm on the host CPU cores (oracle port)

--config selects a BASELINE.json preset (explicit --dtype / --batch / --N / --micro-batch still override it):
  1  configs[1]: batch 32, N = 30, fp32 storage + TF32 MMA                      (the default workload)
  2  configs[2]: batch 256 (4 micro-batches of 64), N = 30, bf16 score net + fp32 SDE state
  3  configs[3]: batch 64, N = 60 ("quality mode"), fp32 storage + TF32 MMA
  4  configs[4]: batch 256 PER GPU (micro-batches of 64), N = 30, bf16: 2048 clips on 8 GPUs, 1/2/4/8-GPU weak scaling

N > 1: launched by torchrun, one rank per GPU; clips shard across ranks with no collective inside the loop and ONE
NCCL all-gather of the enhanced waveforms at the end of each step ("scaling": "weak", per-GPU batch fixed).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_CLIP_EVAL = 2.666e12  # NCSNppLarge on 512 x 640, per clip per evaluation
CLIP_SECONDS = 4.0
L_SAMPLES = 96000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4], help="BASELINE.json configs[i] preset")
    ap.add_argument("--dtype", default=None, choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU per step")
    ap.add_argument("--N", type=int, default=None, help="reverse-diffusion steps")
    ap.add_argument("--micro-batch", type=int, default=None)
    ap.add_argument("--cpu-sample-evals", type=int, default=2, help="network evaluations timed for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed steps of the host-buffer arm (default min(steps, 3))")
    args = ap.parse_args()
    preset = {1: ("fp32", 32, 30, 0), 2: ("bf16", 256, 30, 64), 3: ("fp32", 64, 60, 0), 4: ("bf16", 256, 30, 64)}[args.config]
    for key, val in zip(("dtype", "batch", "N", "micro_batch"), preset):
        if getattr(args, key) is None:
            setattr(args, key, val)
    return args


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            try:
                power.append(float(f[3]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # the sampler only runs during the timed region: the plain median is the clock under load (under the 1 kW cap the
        # loaded clock is LOWER than the idle one, so "top half" would pick the wrong samples)
        load = sm
        power.sort()
        pw = power
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons), "power_w": pw[len(pw) // 2] if pw else None}


def measure_tf32_peak(torch, dev, seconds: float = 4.0, n: int = 8192) -> float:
    """Sustained cuBLAS TF32 GEMM rate (TFLOP/s) on this GPU, measured like MEASURED_PEAKS.json's bf16 figure."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0, iters = time.time(), 0
        e0.record()
        while time.time() - t0 < seconds:
            for _ in range(20):
                torch.matmul(a, b, out=c)
            iters += 20
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        return 2.0 * n ** 3 * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def cpu_oracle_rate(n_evals: int, threads: int):
    """The reference algorithm (oracle port, same op sequence as ScoreModel.sample) on the host cores: one 4 s clip,
    `n_evals` reverse-diffusion steps of the full-size network, extrapolated linearly to N steps."""
    import torch

    from oracle import sgmse_oracle as O

    torch.set_num_threads(threads)
    sd = O.make_state_dict(O.LARGE, seed=7)
    y = O.synthetic_clips(1, L_SAMPLES)
    t0 = time.time()
    O.sample(sd, y, n_evals, seed=42)
    return time.time() - t0


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  /root/reference (pure Python) does not
    travel to the GPU box, so the arm times the oracle port -- bit-exact against the reference (oracle/make_golden.py) --
    with all host threads.  Each step = a bounded sample: one 4 s clip x `cpu_sample_evals` evaluations, scaled to N."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    times = []
    for i in range(args.warmup + args.steps):
        dt = cpu_oracle_rate(args.cpu_sample_evals, threads)
        if i >= args.warmup:
            times.append(dt)
        if sum(times) > 240:  # stay within a few minutes
            break
    per_eval = (sum(times) / len(times)) / args.cpu_sample_evals
    sec_per_clip = per_eval * args.N
    value = 1.0 / sec_per_clip
    sample = f"1 clip x {args.cpu_sample_evals} of N={args.N} steps at full size (512x640), linear in steps"
    line = {
        "impl": "reference", "metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "rtf": sec_per_clip / CLIP_SECONDS,
        "config": {"workload": f"CPU arm: 1 clip x 4 s @ 24 kHz (512x640), {args.cpu_sample_evals} network evaluations timed "
                               f"and extrapolated linearly to N={args.N} PC steps (reverse_diffusion/none), NCSNppLarge, "
                               f"oracle port of the reference, no batch; compared with the GPU arm's batch={args.batch}",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import use_b200
    from oracle import sgmse_oracle as O  # weights + synthetic clips generator only (seeded, shared with the tests)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, N = args.batch, args.N
    model = use_b200.ScoreModel(backbone="ncsnpplarge", sde="ouve", t_eps=3e-2, condition="noisy", sde_input="noisy",
                                n_fft=1022, hop_length=160, num_frames=512, dtype=args.dtype,
                                micro_batch=args.micro_batch or None, N=N)
    model.score_net.load_state_dict(O.make_state_dict(O.LARGE, seed=7), strict=True)
    module = use_b200.SGMSEModule(Score=model)

    # synthetic clips: ONE global batch of world * B clips; use_b200.distributed.sample_sharded gives rank r the contiguous
    # shard [r*B, (r+1)*B) (weights replicated, no collective inside the loop) and all-gathers the enhanced waveforms once
    from use_b200.distributed import sample_sharded, shard_range

    y_all = O.synthetic_clips(B * world, L_SAMPLES)
    lo, hi = shard_range(B * world, rank, world)
    y_host = y_all[lo:hi].contiguous().pin_memory()   # e2e: this rank's clips start in pinned HOST memory every step
    y_dev_all = y_all.to(dev)                          # device-timed arm: inputs already resident in HBM

    def step_device(i):
        return sample_sharded(lambda yl, clip0: model.sample({"perturbed": yl}, N=N, seed=1000 + i, clip0=clip0,
                                                             job_clips=y_dev_all.shape[0])["enhanced"], y_dev_all)

    def step_e2e(i):
        # host -> device copy of this rank's clips, the reference-facing call, gather, device -> host read of the result
        def local(_, clip0):
            batch = {"perturbed": y_host.to(dev, non_blocking=True)}
            return module.predict_step(batch, i, write=False)["enhanced"]

        if world > 1:
            out = sample_sharded(local, y_dev_all)
            return out.to("cpu") if rank == 0 else out[:1].cpu()
        return local(None, 0).cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = model.score_net.engine(dev, args.dtype)

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        l0 = eng.L.use_engine_launch_count(eng.h)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = eng.L.use_engine_launch_count(eng.h) - l0
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    clocks = ClockSampler(local)
    clocks.start()
    ms_total, launches = timed(step_device, args.steps, args.warmup)
    clk = clocks.stop()
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)

    # end to end through the reference-facing call (SGMSEModule.predict_step) with HOST buffers
    e2e_steps = args.e2e_steps if args.e2e_steps > 0 else max(1, min(args.steps, 3))
    ms_e2e, _ = timed(step_e2e, e2e_steps, 1)
    e2e_value = world * B / (ms_e2e / e2e_steps / 1e3)

    # roofline of the dominant kernel class (tcgen05 convolutions): CUDA events around every launch of ONE network
    # evaluation pass (profiling mode, outside the timed region), on the launch stream
    prof = None
    if rank == 0:
        eng.L.use_engine_set_profiling(eng.h, 1)
        model.sample({"perturbed": y_dev_all[lo:hi]}, N=1, seed=7)
        torch.cuda.synchronize()
        import ctypes

        buf = ctypes.create_string_buffer(4096)
        eng.L.use_engine_get_profile(eng.h, buf, 4096)
        prof = json.loads(buf.value.decode())
        big = ctypes.create_string_buffer(1 << 18)
        eng.L.use_engine_get_profile_ops(eng.h, big, 1 << 18)
        eng.L.use_engine_set_profiling(eng.h, 0)
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"profile_ops_{args.dtype}_b{B}.csv"), "w") as f:
                f.write("tag,ms,flops,bytes\n" + big.value.decode())
        except OSError:
            pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
    if args.dtype == "fp32":
        # no TF32 entry in MEASURED_PEAKS.json: measure it here, outside the timed region, with the file's own recipe
        # (cuBLAS GEMM 8192^3 back to back for 4 s, CUDA events) -- a library GEMM as the DENOMINATOR only
        peak = measure_tf32_peak(torch, dev)
        peak_note = (f"measured in-run: torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS TF32), back to back for 4 s, CUDA "
                     f"events = {peak:.1f} TFLOP/s sustained (same recipe as MEASURED_PEAKS.json's bf16 {bf16_peak:.1f})")
    else:
        peak, peak_note = bf16_peak, peak_src
    conv = prof["conv_tc"]
    achieved = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
    net_ms = sum(v["ms"] for k, v in prof.items() if k != "top_conv")
    # DRAM traffic of the conv launches from the committed ncu capture (taken at a small batch; scales with the batch)
    traffic, traffic_note = None, "no ncu summary committed for this dtype"
    try:
        summ_path = next(p for p in (os.path.join(ROOT, "profiles", f"r0{r}_kernel_summary_{args.dtype}.json") for r in (2, 1))
                         if os.path.exists(p))
        summ = json.load(open(summ_path))
        tot_b = sum(v["dram_bytes_total"] for k, v in summ["kernels"].items() if "conv_tc_kernel" in k)
        tot_n = sum(v["launches"] for k, v in summ["kernels"].items() if "conv_tc_kernel" in k)
        cap_b = summ.get("batch", 2)
        if tot_n:
            traffic = tot_b / tot_n / cap_b * B
            traffic_note = (f"dram__bytes_read+write per conv_tc launch from profiles/{os.path.basename(summ_path)} "
                            f"(ncu, batch {cap_b}, {tot_n} launches of one evaluation), scaled linearly to batch {B}")
    except Exception:
        pass
    # clock-independent utilisation of the same kernel from the committed ncu capture (L0 layers, profiles/README.md)
    ncu_pipe = None
    try:
        import csv

        l0_path = next(p for p in (os.path.join(ROOT, "profiles", f"r0{r}_ncu_conv_L0_{args.dtype}.csv") for r in (2, 1))
                       if os.path.exists(p))
        rows = list(csv.reader(open(l0_path)))
        col = rows[0].index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        vals = [float(r[col]) for r in rows[2:] if r and r[col]]
        ncu_pipe = round(sum(vals) / len(vals), 2) if vals else None
    except Exception:
        pass
    roofline = {
        "bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM 3x3/1x1 convolutions, all launches of one "
                                     "network evaluation)",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
        "traffic": traffic, "traffic_note": traffic_note,
        "algorithmic_bytes_per_launch": conv["bytes"] / max(1, conv["launches"]), "peak_source": peak_note,
        "conv_share_of_network_time": conv["ms"] / net_ms if net_ms else None,
        "launches_per_eval": conv["launches"],
        "per_class_ms_per_eval": {k: round(v["ms"], 3) for k, v in prof.items() if k != "top_conv"},
        "whole_path_tflops": world * B * N * FLOP_PER_CLIP_EVAL / (ms_step * 1e-3) / 1e12,
        "ncu_tensor_pipe_active_pct_L0": ncu_pipe,
    }

    cpu_baseline = None
    if not args.no_cpu_baseline and world >= 1:
        threads = os.cpu_count() or 1
        tcpu = cpu_oracle_rate(args.cpu_sample_evals, threads)
        sec_per_clip = tcpu / args.cpu_sample_evals * N
        cpu_baseline = {"value": 1.0 / sec_per_clip, "unit": "clips/s", "cores": threads, "kind": "port",
                        "rtf": sec_per_clip / CLIP_SECONDS,
                        "sample": f"1 clip x {args.cpu_sample_evals} of N={N} steps at full size, extrapolated linearly "
                                  f"in steps; torch {torch.__version__} CPU, {threads} threads"}

    line = {
        "metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if args.dtype == "fp32" else "bf16", "data": "synthetic",
        "rtf": (ms_step / 1e3) / (world * B * CLIP_SECONDS),
        "config": {"workload": f"batch={B} per GPU x 4 s clips @ 24 kHz (512x640 spectrogram), N={N} PC steps "
                               f"(reverse_diffusion / none), NCSNppLarge, {args.dtype} storage"
                               + (" + TF32 tensor-core convolutions" if args.dtype == "fp32" else ", fp32 SDE state"),
                   "baseline_config": args.config, "N": N, "micro_batch": args.micro_batch or None,
                   "global_batch": world * B, "parallelism": f"dp{world} (clips sharded, one NCCL all-gather per step)",
                   "l2_policy": "inputs_exceed_l2 (activations of one step are GBs; no flush needed)",
                   "weights": "seeded random (no checkpoint offline)"},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": B * L_SAMPLES * 4,
                "d2h_bytes_per_step": (world * B if world > 1 else B) * L_SAMPLES * 4, "steps": e2e_steps,
                "api": "SGMSEModule.predict_step(batch) with pinned host input, enhanced waveforms read back"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
