"""ctypes binding of libasr.so (include/asr.h) plus thin tensor-plumbing wrappers.

PyTorch is used only to own device memory and streams; every computation happens inside the C-ABI
library.  There is no CPU fallback: if libasr.so is missing or CUDA is unavailable the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ASR_LIB") or os.path.join(_PKG, "libasr.so")   # ASR_LIB: an experimental build (_build.build(defines=...))

OPTIMIZERS = {"adam": 0, "sgd": 1, "adagrad": 2, "adadelta": 3, "adamax": 4}
OPM_MODES = {"argmax": 0, "slice": 1, "slice_max": 2}
INTERP = {"nearest": 0, "bilinear": 1}


class AsrError(RuntimeError):
    pass


class AsrSolveParams(C.Structure):
    """Mirror of struct AsrSolveParams (include/asr.h)."""
    _fields_ = [
        ("lambda_df", C.c_float), ("lambda_tv", C.c_float), ("lambda_l2", C.c_float), ("lambda_l1", C.c_float),
        ("num_iter", C.c_int32), ("optimizer", C.c_int32),
        ("learning_rate", C.c_float), ("epsilon", C.c_float), ("beta_1", C.c_float), ("beta_2", C.c_float),
        ("amsgrad", C.c_int32),
        ("initial_accumulator_value", C.c_float), ("momentum", C.c_float),
        ("nesterov", C.c_int32), ("lr_scheduler", C.c_int32),
        ("decay_steps", C.c_float), ("decay_rate", C.c_float),
        ("step_offset", C.c_int64),
        ("use_btv", C.c_int32), ("images_in_flight", C.c_int32),
    ]


_fp = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_lib: Optional[C.CDLL] = None

# name -> (restype, argtypes): every symbol include/asr.h declares
SIGNATURES = {
    "asr_version": (C.c_int, []),
    "asr_last_error": (C.c_char_p, []),
    "asr_kernel_launches": (C.c_longlong, []),
    "asr_profile_enable": (C.c_int, [C.c_int]),
    "asr_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "asr_l2_read_probe": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "asr_solve_workspace_bytes": (C.c_int, [C.c_int] * 7 + [C.POINTER(C.c_size_t)]),
    "asr_solve_batched": (C.c_int, [C.POINTER(AsrSolveParams), C.c_int, C.c_void_p, _fp, _fp, _u8p,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "asr_solve_batched_traced": (C.c_int, [C.POINTER(AsrSolveParams), C.c_int, C.c_void_p, _fp, _fp, _u8p,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "asr_solve_sweep": (C.c_int, [C.POINTER(AsrSolveParams), C.c_int, C.c_void_p, _fp, _fp, C.POINTER(C.c_int32),
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "asr_loss_grad_batched": (C.c_int, [C.POINTER(AsrSolveParams), C.c_int, C.c_void_p, C.c_void_p, _fp, _fp, _u8p,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "asr_backproject_batched": (C.c_int, [C.c_int, C.c_void_p, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.c_void_p]),
    "asr_warp_affine": (C.c_int, [C.c_void_p, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p]),
    "asr_backproject_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "asr_backproject_batched_ws": (C.c_int, [C.c_int, C.c_void_p, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "asr_warp_affine_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "asr_warp_affine_ws": (C.c_int, [C.c_void_p, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "asr_opm_extract": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "asr_minmax_normalize": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "asr_threshold": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p]),
    "asr_iou_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "asr_solve_batched_dlpack": (C.c_int, [C.POINTER(AsrSolveParams), C.c_int, C.c_void_p, _fp, _fp, _u8p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}


def lib() -> C.CDLL:
    """Load libasr.so; fail loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AsrError(f"{LIB_PATH} is missing: build it with `python -m "
                           f"deeplabv3plus_augmented_superresolution_b200._build` (needs nvcc); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        missing = [n for n in SIGNATURES if not hasattr(L, n)]
        if missing:
            raise AsrError(f"{LIB_PATH} does not export {missing}: stale build, re-run _build")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(code: int) -> None:
    if code != 0:
        msg = lib().asr_last_error().decode("utf-8", "replace")
        if code == -3:
            raise NotImplementedError(msg)
        raise AsrError(f"libasr error {code}: {msg}")


@dataclass
class SolveParams:
    """Superresolution(...) + Optimizer(...) kwargs as one flat record; defaults = test_SR.py:35-48."""
    lambda_df: float = 1.0
    lambda_tv: float = 0.3
    lambda_l2: float = 0.7
    lambda_l1: float = 0.0
    num_iter: int = 300
    optimizer: str = "adam"
    learning_rate: float = 1e-3
    epsilon: float = 1e-7
    beta_1: float = 0.9
    beta_2: float = 0.999
    amsgrad: bool = True
    initial_accumulator_value: float = 0.1
    momentum: float = 0.0
    nesterov: bool = False
    lr_scheduler: bool = True
    decay_steps: float = 60
    decay_rate: float = 0.3
    step_offset: int = 0
    use_btv: bool = False
    images_in_flight: int = 0

    def to_c(self) -> AsrSolveParams:
        return AsrSolveParams(self.lambda_df, self.lambda_tv, self.lambda_l2, self.lambda_l1, int(self.num_iter),
                              OPTIMIZERS[self.optimizer], self.learning_rate, self.epsilon, self.beta_1, self.beta_2,
                              int(bool(self.amsgrad)), self.initial_accumulator_value, self.momentum,
                              int(bool(self.nesterov)), int(bool(self.lr_scheduler)), float(self.decay_steps),
                              float(self.decay_rate), int(self.step_offset), int(bool(self.use_btv)),
                              int(self.images_in_flight))


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise AsrError("CUDA device required: libasr has no CPU path")
    return torch


def _params_array(params) -> tuple:
    plist = list(params) if isinstance(params, (list, tuple)) else [params]
    arr = (AsrSolveParams * len(plist))(*[p.to_c() for p in plist])
    return arr, len(plist)


def _host_f32(a, shape) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(shape))
    return a


def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Workspace:
    """Caller-owned device scratch for the solve (the library allocates nothing)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device):
        torch = _torch()
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = None
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self.buf


_default_ws = Workspace()
_aux_ws = Workspace()      # scratch of the warp / back-projection calls (kept apart from the solve's, which may be in use on the stream)


def solve_batched(copies, angles, shifts, params, keep=None, want_loss: bool = False, workspace: Optional[Workspace] = None,
                  output_size=None, loss_every: int = 0):
    """asr_solve_batched.  copies: CUDA float32 tensor [B,N,h,w]; angles [B,N], shifts [B,N,2] host arrays;
    params: SolveParams or a list of B of them; output_size (H, W) defaults to (4h, 4w), the reference callers' shape (any even
    integer ratio is accepted); loss_every > 0 also returns the verbose trace [B, ceil(max_iter/loss_every)] (asr_solve_batched_traced).
    Returns x [B,H,W] (CUDA) and, if asked, loss [B] (CUDA), then the trace."""
    torch = _torch()
    L = lib()
    assert copies.is_cuda and copies.dtype == torch.float32 and copies.is_contiguous() and copies.dim() == 4
    B, N, h, w = copies.shape
    H, W = (4 * h, 4 * w) if output_size is None else (int(output_size[0]), int(output_size[1]))
    ang = _host_f32(angles, (B, N))
    shf = _host_f32(shifts, (B, N, 2))
    kp = None if keep is None else np.ascontiguousarray(np.asarray(keep, dtype=np.uint8).reshape(B, N))
    arr, n = _params_array(params)
    max_iter = max(int(p.num_iter) for p in (params if isinstance(params, (list, tuple)) else [params]))
    need = C.c_size_t()
    check(L.asr_solve_workspace_bytes(B, N, h, w, H, W, max_iter, C.byref(need)))
    ws = (workspace or _default_ws).get(need.value, copies.device)
    x = torch.empty((B, H, W), dtype=torch.float32, device=copies.device)
    loss = torch.empty((B,), dtype=torch.float32, device=copies.device) if want_loss else None
    if loss_every > 0:
        cols = -(-max_iter // int(loss_every))
        trace = torch.full((B, max(cols, 1)), float("nan"), dtype=torch.float32, device=copies.device)
        with torch.cuda.device(copies.device):
            check(L.asr_solve_batched_traced(arr, n, copies.data_ptr(), ang.ctypes.data_as(_fp), shf.ctypes.data_as(_fp),
                                             None if kp is None else kp.ctypes.data_as(_u8p), B, N, h, w, H, W,
                                             x.data_ptr(), None if loss is None else loss.data_ptr(), int(loss_every),
                                             trace.data_ptr(), trace.shape[1], ws.data_ptr(), ws.numel(), _stream_ptr(torch)))
        return (x, loss, trace) if want_loss else (x, trace)
    with torch.cuda.device(copies.device):
        check(L.asr_solve_batched(arr, n, copies.data_ptr(), ang.ctypes.data_as(_fp), shf.ctypes.data_as(_fp),
                                  None if kp is None else kp.ctypes.data_as(_u8p), B, N, h, w, H, W,
                                  x.data_ptr(), None if loss is None else loss.data_ptr(),
                                  ws.data_ptr(), ws.numel(), _stream_ptr(torch)))
    return (x, loss) if want_loss else x


def loss_grad_batched(x, copies, angles, shifts, params, keep=None, workspace: Optional[Workspace] = None):
    """asr_loss_grad_batched: one evaluation of residual, gradient and loss at x [B,H,W] (H, W taken from x)."""
    torch = _torch()
    L = lib()
    B, N, h, w = copies.shape
    H, W = int(x.shape[1]), int(x.shape[2])
    assert x.shape[0] == B and x.is_cuda and x.is_contiguous() and copies.is_contiguous()
    ang = _host_f32(angles, (B, N))
    shf = _host_f32(shifts, (B, N, 2))
    kp = None if keep is None else np.ascontiguousarray(np.asarray(keep, dtype=np.uint8).reshape(B, N))
    arr, n = _params_array(params)
    need = C.c_size_t()
    check(L.asr_solve_workspace_bytes(B, N, h, w, H, W, 1, C.byref(need)))
    ws = (workspace or _default_ws).get(need.value, copies.device)
    resid = torch.empty((B, N, h, w), dtype=torch.float32, device=copies.device)
    grad = torch.empty((B, H, W), dtype=torch.float32, device=copies.device)
    loss = torch.empty((B,), dtype=torch.float32, device=copies.device)
    with torch.cuda.device(copies.device):
        check(L.asr_loss_grad_batched(arr, n, x.data_ptr(), copies.data_ptr(), ang.ctypes.data_as(_fp),
                                      shf.ctypes.data_as(_fp), None if kp is None else kp.ctypes.data_as(_u8p),
                                      B, N, h, w, H, W, resid.data_ptr(), grad.data_ptr(), loss.data_ptr(),
                                      ws.data_ptr(), ws.numel(), _stream_ptr(torch)))
    return resid, grad, loss


def solve_sweep(copies, angles, shifts, params_list, stack_index, want_loss: bool = False, workspace: Optional[Workspace] = None):
    """asr_solve_sweep: len(params_list) solves, point i on stack stack_index[i] of copies [S,N,h,w]."""
    torch = _torch()
    L = lib()
    assert copies.is_cuda and copies.dtype == torch.float32 and copies.is_contiguous() and copies.dim() == 4
    S, N, h, w = copies.shape
    H, W = 4 * h, 4 * w
    P = len(params_list)
    ang = _host_f32(angles, (S, N))
    shf = _host_f32(shifts, (S, N, 2))
    idx = np.ascontiguousarray(np.asarray(stack_index, dtype=np.int32).reshape(P))
    arr, n = _params_array(list(params_list))
    need = C.c_size_t()
    check(L.asr_solve_workspace_bytes(P, N, h, w, H, W, max(int(p.num_iter) for p in params_list), C.byref(need)))
    ws = (workspace or _default_ws).get(need.value, copies.device)
    x = torch.empty((P, H, W), dtype=torch.float32, device=copies.device)
    loss = torch.empty((P,), dtype=torch.float32, device=copies.device) if want_loss else None
    with torch.cuda.device(copies.device):
        check(L.asr_solve_sweep(arr, n, copies.data_ptr(), ang.ctypes.data_as(_fp), shf.ctypes.data_as(_fp),
                                idx.ctypes.data_as(C.POINTER(C.c_int32)), S, N, h, w, H, W, x.data_ptr(),
                                None if loss is None else loss.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(torch)))
    return (x, loss) if want_loss else x
