"""ctypes binding of libbclinear.so (include/bclinear.h).  No torch types cross this boundary: only raw
pointers, sizes and enums.  There is no fallback: if the library is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BCL_LIB_PATH") or os.path.join(_HERE, "libbclinear.so")   # override: A/B builds when profiling

ABI_VERSION = 14
F32, F16, BF16 = 0, 1, 2
LAYOUT_STREAM, LAYOUT_GEMM, LAYOUT_DECODE_TC = 0, 1, 2
EPILOGUE_LINEAR, EPILOGUE_SWIGLU, EPILOGUE_EMIT_SPECTRUM = 0, 1, 2
X_SPECTRUM = 3                      # x_dtype of a chain problem that consumes an emitted spectrum
CHAIN_MAX_PROBLEMS, CHAIN_MAX_STEPS = 8, 8

DTYPE = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}

# every symbol include/bclinear.h declares (tests check the built library exports all of them)
SYMBOLS = (
    "bcl_abi_version", "bcl_last_error", "bcl_spectrum_bytes", "bcl_precompute_spectrum", "bcl_workspace_bytes",
    "bcl_forward", "bcl_group_plan_bytes", "bcl_group_plan_encode", "bcl_forward_grouped", "bcl_forward_host", "bcl_forward_grouped_host", "bcl_forward_grouped_host_pipelined",
    "bcl_project_from_dense",
    "bcl_forward_bcast", "bcl_conj_transpose", "bcl_backward_workspace_bytes", "bcl_backward_weights", "bcl_prefill_geometry", "bcl_prefill_rfft", "bcl_prefill_gemm", "bcl_prefill_irfft",
    "bcl_launch_noop", "bcl_forward_chain", "bcl_push_columns", "bcl_forward_swiglu", "bcl_group_plan_encode_layout", "bcl_forward_grouped_host_async", "bcl_group_plan_bytes_for",
)


class BclError(RuntimeError):
    pass


class Problem(ctypes.Structure):
    """struct bcl_problem"""
    _fields_ = [("x", ctypes.c_void_p), ("spec", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("y", ctypes.c_void_p),
                ("N", ctypes.c_int64), ("x_row_stride", ctypes.c_int64), ("y_row_stride", ctypes.c_int64),
                ("p", ctypes.c_int32), ("q", ctypes.c_int32),
                ("x_dtype", ctypes.c_int32), ("y_dtype", ctypes.c_int32), ("bias_dtype", ctypes.c_int32),
                ("epilogue", ctypes.c_int32), ("y_spectrum", ctypes.c_void_p)]


_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) and type the library.  Raises loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built.  Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C llama-fft_b200/csrc`.  "
            "There is no CPU or torch.fft fallback for the BC-linear hot path.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t
    L.bcl_abi_version.restype = i32
    L.bcl_last_error.restype = ctypes.c_char_p
    L.bcl_spectrum_bytes.restype = sz
    L.bcl_spectrum_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.bcl_precompute_spectrum.argtypes = [vp, i32, i32, i32, vp, i32, i32, vp]
    L.bcl_workspace_bytes.restype = sz
    L.bcl_workspace_bytes.argtypes = [i64, i32, i32, i32, i32, i32]
    L.bcl_forward.argtypes = [vp, i32, i64, i64, vp, i32, i32, i32, i32, i32, vp, i32, vp, i32, i64, vp, sz, vp]
    L.bcl_forward_bcast.argtypes = [vp, i32, i64, i64, vp, i32, i32, i32, i32, i32, vp, i32, ctypes.POINTER(vp), i32, i32, i64, vp, sz, vp]
    L.bcl_conj_transpose.argtypes = [vp, i32, i32, i32, vp, vp]
    L.bcl_backward_workspace_bytes.restype = sz
    L.bcl_backward_workspace_bytes.argtypes = [i64, i32, i32, i32]
    L.bcl_backward_weights.argtypes = [vp, i32, i64, vp, i32, i64, i64, i32, i32, i32, vp, i32, vp, sz, vp]
    L.bcl_group_plan_bytes.restype = sz
    L.bcl_group_plan_bytes.argtypes = [i32, i32]
    L.bcl_group_plan_bytes_for.restype = sz
    L.bcl_group_plan_bytes_for.argtypes = [ctypes.POINTER(Problem), i32, i32, i32, i32]
    L.bcl_group_plan_encode.argtypes = [ctypes.POINTER(Problem), i32, i32, i32, i32, vp, sz]
    L.bcl_group_plan_encode_layout.argtypes = [ctypes.POINTER(Problem), i32, i32, i32, i32, i32, vp, sz]
    L.bcl_forward_grouped.argtypes = [vp, vp, vp]
    L.bcl_forward_host.argtypes = [vp, i32, i64, vp, i32, i32, i32, i32, i32, vp, i32, vp, i32, vp, vp, vp, sz, vp]
    L.bcl_forward_grouped_host.argtypes = [vp, vp, vp, vp, sz, vp, vp, sz, vp]
    L.bcl_forward_grouped_host_async.argtypes = [vp, vp, vp, vp, sz, vp, vp, sz, vp]
    L.bcl_forward_grouped_host_pipelined.argtypes = [vp, vp, vp, vp, sz, vp, vp, sz, vp, vp, vp, vp, vp, vp, i32]
    L.bcl_project_from_dense.argtypes = [vp, i32, i64, i32, i32, i32, vp, vp]
    ip = ctypes.POINTER(ctypes.c_int)
    L.bcl_prefill_geometry.restype = None
    szp = ctypes.POINTER(ctypes.c_size_t)
    L.bcl_prefill_geometry.argtypes = [i32, i32, i32, ip, ip, ip, szp, szp, szp]
    L.bcl_prefill_rfft.argtypes = [vp, i32, i64, i64, i32, i32, i32, vp, i32, vp, vp, vp]
    L.bcl_prefill_gemm.argtypes = [vp, vp, vp, i64, i32, i32, i32, vp, vp]
    L.bcl_prefill_irfft.argtypes = [vp, i64, i32, i32, i32, vp, i32, i32, vp, i32, vp, i32, i64, vp, vp]
    L.bcl_launch_noop.argtypes = [i32, vp]
    L.bcl_forward_swiglu.argtypes = [vp, i32, i64, i64, vp, i32, i32, i32, i32, vp, i32, i64, vp, sz, vp]
    L.bcl_push_columns.argtypes = [vp, ctypes.POINTER(vp), i32, i64, i64, i64, i64, i32, vp]
    L.bcl_forward_chain.argtypes = [ctypes.POINTER(Problem), ctypes.POINTER(ctypes.c_int), i32, i32, i32, vp, vp]
    for name in SYMBOLS:
        getattr(L, name)
    if L.bcl_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH} has ABI {L.bcl_abi_version()}, this package needs {ABI_VERSION}: rebuild it")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    """Map a bcl_status to the exception the reference's Python would have raised."""
    if rc == 0:
        return
    msg = lib().bcl_last_error().decode("utf-8", "replace")
    text = f"{what}: {msg} (bcl_status {rc})"
    if rc in (-1, -2, -3, -5):          # ARG, SHAPE, BLOCK, DTYPE
        raise ValueError(text)
    if rc == -8:
        raise NotImplementedError(text)
    raise BclError(text)


def prefill_geometry(p: int, q: int, B: int):
    """(F, Kp, Np) of the GEMM layout: B/2 frequency slots, 2q rounded up to 16, 2p rounded up to 64."""
    F, Kp, Np = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib().bcl_prefill_geometry(p, q, B, ctypes.byref(F), ctypes.byref(Kp), ctypes.byref(Np), None, None, None)
    return F.value, Kp.value, Np.value


def prefill_dft_offsets(p: int, q: int, B: int):
    """Byte offsets of the (fp16 forward, bf16 forward, bf16 inverse) DFT matrices inside the GEMM-layout spectrum buffer."""
    a, b, c = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    lib().bcl_prefill_geometry(p, q, B, None, None, None, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return a.value, b.value, c.value


def stream_ptr(device) -> int:
    """Raw cudaStream_t of torch's current stream on `device` (the private accessor skips the Stream object: ~0.5 us instead of ~6 us
    per call, which matters when a decode step makes hundreds of per-module calls)."""
    try:
        dev = device if isinstance(device, torch.device) else torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        return torch._C._cuda_getCurrentRawStream(idx)
    except AttributeError:                                    # accessor renamed in another torch version: the public route
        return torch.cuda.current_stream(device).cuda_stream


class on_device:
    """`with on_device(t.device):` -- torch.cuda.device(...) only when the current device differs (the context manager costs ~4 us)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self.ctx = None if idx == torch.cuda.current_device() else torch.cuda.device(idx)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} is on {t.device}: the BC-linear forward runs only as sm_100a CUDA kernels "
            "(no CPU / torch.fft fallback).  Move the module and its input to a CUDA device.")
