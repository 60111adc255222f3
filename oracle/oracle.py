"""ctypes front-end of the CPU oracle (oracle/asr_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see asr_oracle.h): this is a restatement
of the reference's TensorFlow arithmetic, not the reference itself.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_OPT = {"adam": 0, "sgd": 1, "adagrad": 2, "adadelta": 3, "adamax": 4}
_OPM = {"argmax": 0, "slice": 1, "slice_max": 2}


class _Params(C.Structure):
    _fields_ = [
        ("lambda_df", C.c_float), ("lambda_tv", C.c_float), ("lambda_l2", C.c_float), ("lambda_l1", C.c_float),
        ("num_iter", C.c_int32), ("optimizer", C.c_int32),
        ("learning_rate", C.c_float), ("epsilon", C.c_float), ("beta_1", C.c_float), ("beta_2", C.c_float),
        ("amsgrad", C.c_int32),
        ("initial_accumulator_value", C.c_float), ("momentum", C.c_float),
        ("nesterov", C.c_int32), ("lr_scheduler", C.c_int32),
        ("decay_steps", C.c_float), ("decay_rate", C.c_float),
        ("step_offset", C.c_int64),
        ("use_btv", C.c_int32), ("n_keep", C.c_int32),
    ]


@dataclass
class SolveParams:
    """kwargs of Superresolution.__init__ (superresolution.py:26-42) + Optimizer.__init__ (optimizer.py:4-48).
    Defaults are the test_SR.py constants (test_SR.py:35-48)."""
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

    def to_c(self) -> _Params:
        return _Params(self.lambda_df, self.lambda_tv, self.lambda_l2, self.lambda_l1, self.num_iter,
                       _OPT[self.optimizer], self.learning_rate, self.epsilon, self.beta_1, self.beta_2,
                       int(self.amsgrad), self.initial_accumulator_value, self.momentum, int(self.nesterov),
                       int(bool(self.lr_scheduler)), float(self.decay_steps), float(self.decay_rate),
                       int(self.step_offset), int(self.use_btv), 0)


def build(variant: str = "") -> str:
    """Compile oracle/_build/liboracle{variant}.so with the committed Makefile (gcc, -ffp-contract=off)."""
    name = f"liboracle{variant}.so"
    target = os.path.join(_HERE, "_build", name)
    src = [os.path.join(_HERE, f) for f in ("asr_oracle.c", "asr_oracle.h", "Makefile")]
    if not os.path.exists(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, f"_build/{name}"])
    return target


_libs: dict[str, C.CDLL] = {}
_fp = C.POINTER(C.c_float)


def _f(a):
    return a.ctypes.data_as(_fp)


def lib(variant: str = "") -> C.CDLL:
    if variant not in _libs:
        L = C.CDLL(build(variant))
        L.orc_loss_and_grad.restype = C.c_float
        L.orc_augmented_superresolution.restype = C.c_float
        L.orc_single_class_iou.restype = C.c_double
        _libs[variant] = L
    return _libs[variant]


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return lib().orc_num_threads()


def use_all_cores() -> int:
    """Give the oracle every core this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(int(n))
    return num_threads()


def rotate_matrix(angle, H, W):
    t = np.zeros(8, np.float32)
    lib().orc_rotate_matrix(C.c_float(angle), H, W, _f(t))
    return t


def translate_matrix(dx, dy):
    t = np.zeros(8, np.float32)
    lib().orc_translate_matrix(C.c_float(dx), C.c_float(dy), _f(t))
    return t


def invert_transform(t):
    t = _c32(t)
    o = np.zeros(8, np.float32)
    lib().orc_invert_transform(_f(t), _f(o))
    return o


def projective_transform(images, transforms, interpolation="bilinear", fill_value=0.0):
    """tfa.image.transform -> ImageProjectiveTransformV3.  images [N,H,W,C], transforms [N,8] or [8]."""
    images = _c32(images)
    N, H, W, Cc = images.shape
    tr = _c32(transforms).reshape(-1, 8)
    out = np.empty_like(images)
    lib().orc_projective_transform(_f(images), N, H, W, Cc, _f(tr), tr.shape[0],
                                   1 if interpolation.lower() == "bilinear" else 0, C.c_float(fill_value), _f(out))
    return out


def rotate(images, angles, interpolation="bilinear"):
    images = _c32(images)
    N, H, W, _ = images.shape
    angles = np.asarray(angles, np.float32).reshape(-1)
    if angles.shape[0] == 1 and N > 1:
        angles = np.repeat(angles, N)
    tr = np.stack([rotate_matrix(float(a), H, W) for a in angles])
    return projective_transform(images, tr, interpolation)


def translate(images, shifts, interpolation="bilinear"):
    images = _c32(images)
    N = images.shape[0]
    shifts = np.asarray(shifts, np.float32).reshape(-1, 2)
    if shifts.shape[0] == 1 and N > 1:
        shifts = np.repeat(shifts, N, 0)
    tr = np.stack([translate_matrix(float(s[0]), float(s[1])) for s in shifts])
    return projective_transform(images, tr, interpolation)


def resize_bilinear(images, size):
    images = _c32(images)
    N, h, w, Cc = images.shape
    out = np.empty((N, size[0], size[1], Cc), np.float32)
    lib().orc_resize_bilinear(_f(images), N, h, w, Cc, _f(out), size[0], size[1])
    return out


def resize_bilinear_grad(grad, orig_size):
    grad = _c32(grad)
    N, H, W, Cc = grad.shape
    out = np.empty((N, orig_size[0], orig_size[1], Cc), np.float32)
    lib().orc_resize_bilinear_grad(_f(grad), N, H, W, Cc, _f(out), orig_size[0], orig_size[1])
    return out


def _prep(copies, angles, shifts):
    copies = _c32(copies)
    if copies.ndim == 4:
        copies = copies[..., 0]
    angles = _c32(angles).reshape(-1)
    shifts = _c32(shifts).reshape(-1, 2)
    assert copies.shape[0] == angles.shape[0] == shifts.shape[0]
    return np.ascontiguousarray(copies), angles, np.ascontiguousarray(shifts)


def loss_and_grad(x, copies, angles, shifts, params: SolveParams, keep=None, variant="", want_resid=False):
    """superresolution.py:44-100 + tape.gradient.  x [H,W]; returns (loss, grad[H,W][, resid[N,h,w]])."""
    copies, angles, shifts = _prep(copies, angles, shifts)
    x = _c32(x).reshape(x.shape[0], x.shape[1])
    H, W = x.shape
    N, h, w = copies.shape
    g = np.empty((H, W), np.float32)
    r = np.empty((N, h, w), np.float32) if want_resid else None
    p = params.to_c()
    kp = None if keep is None else np.ascontiguousarray(keep, np.uint8)
    loss = lib(variant).orc_loss_and_grad(_f(x), H, W, _f(copies), N, h, w, _f(angles), _f(shifts), C.byref(p),
                                          None if kp is None else kp.ctypes.data_as(C.POINTER(C.c_uint8)),
                                          _f(g), None if r is None else _f(r))
    return (float(loss), g, r) if want_resid else (float(loss), g)


def augmented_superresolution(copies, angles, shifts, params: SolveParams, output_size=(512, 512), keep=None,
                              trace_iters=(), variant=""):
    """superresolution.py:102-137.  Returns (x [H,W,1] f32, loss float[, trace [T,H,W]])."""
    copies, angles, shifts = _prep(copies, angles, shifts)
    N, h, w = copies.shape
    H, W = output_size
    x = np.empty((H, W), np.float32)
    p = params.to_c()
    kp = None if keep is None else np.ascontiguousarray(keep, np.uint8)
    ti = np.ascontiguousarray(trace_iters, np.int32)
    tr = np.empty((len(ti), H, W), np.float32)
    loss = lib(variant).orc_augmented_superresolution(
        _f(copies), N, h, w, H, W, _f(angles), _f(shifts), C.byref(p),
        None if kp is None else kp.ctypes.data_as(C.POINTER(C.c_uint8)), _f(x),
        ti.ctypes.data_as(C.POINTER(C.c_int32)), len(ti), _f(tr))
    if len(ti):
        return x[..., None], float(loss), tr
    return x[..., None], float(loss)


def backproject(copies, angles, shifts, mode, output_size=(512, 512)):
    """superresolution.py:139-161; mode 'max' | 'mean'.  Returns [H,W,1]."""
    copies, angles, shifts = _prep(copies, angles, shifts)
    N, h, w = copies.shape
    H, W = output_size
    out = np.empty((H, W), np.float32)
    lib().orc_backproject(_f(copies), N, h, w, H, W, _f(angles), _f(shifts), 0 if mode == "max" else 1, _f(out))
    return out[..., None]


def threshold_image(image, th_value, th_factor=0.15, th_mask=None):
    """superres_utils.py:118-139 -> int32, same shape."""
    image = _c32(image)
    out = np.empty(image.shape, np.int32)
    m = None if th_mask is None else _c32(th_mask)
    lib().orc_threshold(_f(image), C.c_int64(image.size), int(th_value), C.c_float(th_factor),
                        None if m is None else _f(m), out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


def minmax_normalize_global(a, new_min=0.0, new_max=1.0):
    a = _c32(a)
    out = np.empty_like(a)
    lib().orc_minmax_normalize_global(_f(a), C.c_int64(a.size), C.c_float(new_min), C.c_float(new_max), _f(out))
    return out


def opm_extract(logits, class_id, mode):
    """augmentation_utils.py:80-115.  logits [N,h,w,K] -> (class [N,h,w,1], max [N,h,w,1] | None)."""
    logits = _c32(logits)
    N, h, w, K = logits.shape
    co = np.empty((N, h, w), np.float32)
    mo = np.empty((N, h, w), np.float32) if mode == "slice_max" else None
    lib().orc_opm_extract(_f(logits), N, h, w, K, int(class_id), _OPM[mode], _f(co), None if mo is None else _f(mo))
    return co[..., None], (None if mo is None else mo[..., None])


def compute_iou(true_image, image, class_id, include_bg=False):
    """utils.py:207-230 single-class branch."""
    t = np.ascontiguousarray(np.asarray(true_image).reshape(-1), np.int32)
    p = np.ascontiguousarray(np.asarray(image).reshape(-1), np.int32)
    ip = C.POINTER(C.c_int32)
    return lib().orc_single_class_iou(t.ctypes.data_as(ip), p.ctypes.data_as(ip), C.c_int64(t.size), int(class_id),
                                      int(include_bg))
