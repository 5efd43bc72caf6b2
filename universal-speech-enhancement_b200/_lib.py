"""ctypes binding of libuse_b200.so (C ABI declared in include/use_b200.h).

The product path has NO fallback: if the library is missing or a call fails, a RuntimeError is raised
(the reference surfaces native-op failures the same way through TORCH_CHECK,
/root/reference/src/models/components/sgmse/backbones/ncsnpp_utils/op/upfirdn2d.cpp:8-10).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuse_b200.so")
CSRC = os.path.join(_HERE, "csrc")

DTYPE_F32 = 0
DTYPE_BF16 = 1
DTYPE_F32X3 = 2  # engine only: 3xTF32 split convolutions (fp32-level accuracy; parity mode)

PRED = {"reverse_diffusion": 0, "euler_maruyama": 1, "none": 2}
CORR = {"none": 0, "langevin": 1, "ald": 2}

# every symbol include/use_b200.h declares (tests check the header against this list and the .so)
SYMBOLS = [
    "use_abi_version", "use_last_error", "use_engine_create", "use_engine_destroy", "use_engine_set_weight",
    "use_engine_pack", "use_engine_upload", "use_engine_workspace_bytes", "use_engine_set_option", "use_engine_launch_count", "use_engine_set_profiling",
    "use_engine_get_profile", "use_engine_get_profile_ops", "use_score_forward", "use_score_forward2", "use_reverse_drift", "use_net_forward", "use_pc_sample", "use_pc_sample_ex", "use_train_forward",
    "use_stft", "use_istft", "use_resample_workspace_bytes", "use_resample_fft_f32", "use_peak_normalize_pad_f32", "use_upfirdn2d_f32", "use_op_gn_stats", "use_op_gn_apply", "use_op_conv_tc", "use_op_gn_affine", "use_op_conv_tc_gn", "use_op_head_tc", "use_op_head_tc_gn", "use_op_set_latency", "use_op_combine_stats", "use_op_gn_apply_aff",
    "use_op_conv_ref", "use_op_conv_in4", "use_op_conv_out4", "use_op_combine", "use_op_fir4_down", "use_op_philox",
    "use_pack_conv_weight", "use_pack_head_weight",
]


class UseConfig(C.Structure):
    _fields_ = [
        ("nf", C.c_int), ("num_levels", C.c_int), ("ch_mult", C.c_int * 8), ("num_res_blocks", C.c_int),
        ("input_channels", C.c_int), ("act_dtype", C.c_int), ("n_fft", C.c_int), ("hop", C.c_int),
        ("spec_factor", C.c_float), ("spec_abs_exponent", C.c_float), ("theta", C.c_float),
        ("conditional", C.c_int), ("scale_by_sigma", C.c_int),
    ]


class UseSamplerOpts(C.Structure):
    _fields_ = [
        ("predictor", C.c_int), ("corrector", C.c_int), ("corrector_steps", C.c_int), ("snr", C.c_float),
        ("probability_flow", C.c_int), ("denoise", C.c_int), ("g_host", C.c_void_p), ("ald_step_host", C.c_void_p),
        ("trace", C.c_void_p), ("x_init", C.c_void_p), ("dt_steps", C.c_int),
        ("cond", C.c_void_p), ("cond2", C.c_void_p),
    ]


_lock = threading.Lock()
_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into libuse_b200.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libuse_b200.so failed:\n" + res.stderr[-4000:])
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the library (once).  Raises RuntimeError when it is absent: there is no CPU / torch fallback."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"use_b200: {LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C universal-speech-enhancement_b200/csrc`). There is no fallback path."
            )
        L = C.CDLL(LIB_PATH)
        vp, i32, f32, u64, u32, sz = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_uint32, C.c_size_t
        L.use_abi_version.restype = i32
        L.use_last_error.restype = C.c_char_p
        L.use_engine_create.restype = vp
        L.use_engine_create.argtypes = [C.POINTER(UseConfig)]
        L.use_engine_destroy.restype = None
        L.use_engine_destroy.argtypes = [vp]
        L.use_engine_set_weight.argtypes = [vp, C.c_char_p, vp, C.POINTER(C.c_int64), i32]
        L.use_engine_pack.argtypes = [vp, C.POINTER(sz)]
        L.use_engine_upload.argtypes = [vp, vp, sz, vp]
        L.use_engine_workspace_bytes.argtypes = [vp, i32, i32, i32, C.POINTER(sz)]
        L.use_engine_set_option.argtypes = [vp, C.c_char_p, i32]
        L.use_engine_launch_count.restype = C.c_longlong
        L.use_engine_launch_count.argtypes = [vp]
        L.use_engine_set_profiling.argtypes = [vp, i32]
        L.use_engine_get_profile.argtypes = [vp, C.c_char_p, sz]
        L.use_engine_get_profile_ops.argtypes = [vp, C.c_char_p, sz]
        L.use_score_forward.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp]
        L.use_score_forward2.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp]
        L.use_reverse_drift.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, f32, i32, vp, vp, sz, vp]
        L.use_net_forward.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp]
        L.use_pc_sample.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, f32, vp, u64, u32, vp, sz, vp]
        L.use_pc_sample_ex.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, f32, vp, u64, u32,
                                       C.POINTER(UseSamplerOpts), vp, sz, vp]
        L.use_train_forward.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, u64, u32, i32, vp, vp, vp, sz, vp]
        L.use_stft.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp]
        L.use_istft.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
        L.use_resample_workspace_bytes.argtypes = [i32, i32, i32, C.POINTER(sz)]
        L.use_resample_fft_f32.argtypes = [vp, i32, i32, vp, i32, i32, vp, sz, vp]
        L.use_peak_normalize_pad_f32.argtypes = [vp, vp, i32, i32, f32, vp, vp]
        L.use_upfirdn2d_f32.argtypes = [vp, vp, i32, i32, i32, i32, vp] + [i32] * 10 + [vp]
        L.use_op_gn_stats.argtypes = [i32, vp, vp, i32, i32, i32, vp]
        L.use_op_gn_apply.argtypes = [i32, vp, vp, i32, vp, vp, i32, vp, vp, f32, i32, i32, i32, vp, vp, i32, i32, i32, vp]
        L.use_op_gn_apply_aff.argtypes = [i32, vp, vp, i32, vp, vp, i32, vp, vp, f32, i32, i32, i32, vp, vp, i32, i32, i32, vp, vp]
        L.use_op_conv_tc.argtypes = [i32, i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                                     C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), i32, i32, i32, i32,
                                     vp, i32, vp, f32, vp, vp, vp]
        L.use_op_gn_affine.argtypes = [vp, i32, vp, i32, vp, vp, f32, i32, vp, i32, vp]
        L.use_op_conv_tc_gn.argtypes = [i32, i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                                        C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(vp),
                                        C.POINTER(i32), C.POINTER(i32), i32, i32, i32, i32, vp, i32, vp, f32, vp, vp, vp]
        L.use_op_conv_ref.argtypes = [i32, vp, vp, vp, i32, vp, f32, vp, i32, i32, i32, i32, i32, i32, vp]
        L.use_op_conv_in4.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, vp]
        L.use_op_conv_out4.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
        L.use_op_head_tc.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
        L.use_op_set_latency.argtypes = [i32]
        L.use_op_head_tc_gn.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
        L.use_op_combine.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
        L.use_op_combine_stats.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
        L.use_op_fir4_down.argtypes = [vp, vp, i32, i32, i32, i32, vp]
        L.use_op_philox.argtypes = [vp, u64, u32, u32, i32, sz, vp]
        L.use_pack_conv_weight.argtypes = [i32, vp, i32, i32, i32, vp]
        L.use_pack_head_weight.argtypes = [i32, vp, i32, i32, vp]
        for name in SYMBOLS:
            fn = getattr(L, name)  # AttributeError here = header / library mismatch
            if fn.restype is C.c_int and name not in ("use_abi_version",):
                fn.restype = C.c_int
        if L.use_abi_version() != 5:
            raise RuntimeError("use_b200: ABI version mismatch between the Python layer and libuse_b200.so")
        _lib = L
        return L


def check(rc: int, what: str = "use_b200") -> None:
    if rc != 0:
        msg = lib().use_last_error()
        raise RuntimeError(f"{what} failed: {msg.decode() if msg else rc}")


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def dtype_code(name) -> int:
    """'fp32' / torch.float32 / 0 -> DTYPE_F32 (TF32 MMA);  'bf16' / torch.bfloat16 / 1 -> DTYPE_BF16."""
    if isinstance(name, int) and not isinstance(name, bool) and name in (DTYPE_F32, DTYPE_BF16, DTYPE_F32X3):
        return name
    s = str(name).replace("torch.", "").lower()
    if s in ("fp32", "float32", "tf32", "float"):
        return DTYPE_F32
    if s in ("bf16", "bfloat16"):
        return DTYPE_BF16
    if s in ("fp32x3", "3xtf32", "tf32x3"):
        return DTYPE_F32X3
    raise ValueError(f"unsupported compute dtype {name!r} (use 'fp32', 'bf16' or 'fp32x3')")
