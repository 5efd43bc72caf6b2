"""Shared helpers of the parity tests (torch is the checker here, never the product path)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

F32, BF16 = 0, 1
DTYPES = {"fp32": F32, "bf16": BF16}


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to TF32 (10 mantissa bits), what the kernels / the weight packer apply."""
    i = x.contiguous().view(torch.int32)
    i = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


def to_operand(x: torch.Tensor, dt: int) -> torch.Tensor:
    """fp32 tensor -> the values an MMA operand of act dtype dt holds (as fp32)."""
    return x.to(torch.bfloat16).to(torch.float32) if dt == BF16 else round_tf32(x.to(torch.float32))


def act_tensor(x_nchw: torch.Tensor, dt: int, device="cuda") -> torch.Tensor:
    """NCHW fp32 -> NHWC contiguous device tensor of the act dtype."""
    x = x_nchw.permute(0, 2, 3, 1).contiguous()
    return x.to(device=device, dtype=torch.bfloat16 if dt == BF16 else torch.float32)


def from_act(x_nhwc: torch.Tensor) -> torch.Tensor:
    return x_nhwc.to(torch.float32).permute(0, 3, 1, 2).contiguous().cpu()


def pack_weight(L, w_oihw: torch.Tensor, dt: int, device="cuda") -> torch.Tensor:
    O, I, k, _ = w_oihw.shape
    w = w_oihw.to(torch.float32).contiguous()
    out = torch.empty(k * k * O * I * (2 if dt == BF16 else 4), dtype=torch.uint8)
    rc = L.use_pack_conv_weight(dt, w.data_ptr(), O, I, k, out.data_ptr())
    assert rc == 0
    return out.to(device)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.to(torch.float64), b.to(torch.float64)
    return float((a - b).norm() / (b.norm() + 1e-30))


def describe_mismatch(got: torch.Tensor, ref: torch.Tensor, k: int = 5) -> str:
    d = (got.double() - ref.double()).abs()
    idx = torch.topk(d.flatten(), min(k, d.numel())).indices
    lines = [f"max|d|={float(d.max()):.4g} rel_l2={rel_l2(got, ref):.4g} |ref|max={float(ref.abs().max()):.4g}"]
    for i in idx.tolist():
        pos = np.unravel_index(i, tuple(d.shape))
        lines.append(f"  at {tuple(int(p) for p in pos)}: got {float(got.flatten()[i]):.6g} ref {float(ref.flatten()[i]):.6g}")
    return "\n".join(lines)


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)


def int_array(vals):
    return (C.c_int * len(vals))(*vals)
