"""Batched form of the reference's per-image loop (SR_single_class.py:83-127, sweep_script.py:88-130).

The reference walks a directory of hdf5 augmented-copies files one image at a time: load_SR_data ->
compute_SR("aug") -> compute_SR("max") -> compute_SR("mean") -> six compute_IoU calls.  Driving libasr
through that loop keeps one image in flight (single-image latency, DESIGN.md "Roofline"); this module
does the same work for `batch` images per launch and returns exactly what the sequential loop would:

* the augmented solve of image j starts at optimizer step `iterations + j * num_iter` (one Optimizer is
  shared by the whole run and Keras never resets its counter; slice_max mode consumes two solves per
  image, class map first, then the max map: superres_utils.py:250-254);
* thresholds follow compute_SR: class >= max when the file holds max_masks, else x > th_factor * max(x);
* files that load_SR_data rejects are reported and skipped (SR_single_class.py:85-90).

Everything stays on the device between the hdf5 read and the IoU counts; only file IO and the final
divisions run on the host.  New logic lives here so that the reference-named modules stay plain mirrors.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Callable, Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from .superresolution_scripts.superresolution import Superresolution
from .superresolution_scripts.superres_utils import load_SR_data
from .utils import compute_IoU_batched


@dataclass
class BatchResult:
    """Per-image outputs of one batch, in file order.  Masks are CUDA int32 [B,H,W] with values {0, class_id}."""
    filenames: List[str]
    paths: List[str]
    aug: object = None
    max: object = None
    mean: object = None
    skipped: List[str] = field(default_factory=list)


def _threshold_batched(x, class_id, th_factor, th_mask=None):
    """threshold_image (superres_utils.py:118-139) for B maps at once: CUDA [B,H,W] f32 -> int32."""
    torch = _lib._torch()
    L = _lib.lib()
    B = x.shape[0]
    n = x[0].numel()
    out = torch.empty(x.shape, dtype=torch.int32, device=x.device)
    ws = torch.empty(2 * B, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.asr_threshold(x.data_ptr(), B, n, int(class_id), float(th_factor),
                                   None if th_mask is None else th_mask.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                   _lib._stream_ptr(torch)))
    return out


def solve_class_and_max(sr: Superresolution, class_masks, max_masks, angles, shifts):
    """compute_SR's slice_max branch (superres_utils.py:250-254: the class solve, then the max solve, on one shared optimizer)
    as ONE two-image call: same step offsets, same results, both solves in flight together.  Returns two [H,W,1] float32 arrays."""
    from .superresolution_scripts.superresolution import _as_device_stack
    torch = _lib._torch()
    cls, mx = _as_device_stack(class_masks), _as_device_stack(max_masks)
    ang = np.asarray(angles, np.float32).reshape(1, -1)
    shf = np.asarray(shifts, np.float32).reshape(1, -1, 2)
    keep = sr._dropout_keep(cls.shape[0])
    x = sr.augmented_superresolution_batched(torch.stack([cls, mx]).contiguous(), np.repeat(ang, 2, axis=0), np.repeat(shf, 2, axis=0),
                                             keep=None if keep is None else np.stack([keep, keep]))
    x = x.cpu().numpy()
    return x[0][..., None], x[1][..., None]


def solve_stacks(sr: Superresolution, class_stacks, max_stacks, angles, shifts, class_id, th_factor,
                 sr_types: Sequence[str] = ("aug", "max", "mean")) -> dict:
    """compute_SR for B images at once.  class_stacks: CUDA [B,N,h,w]; max_stacks: same shape or None;
    angles [B,N], shifts [B,N,2].  Returns {sr_type: CUDA int32 [B,H,W]}.  Advances sr.optimizer.iterations
    exactly as B sequential compute_SR("aug") calls would."""
    torch = _lib._torch()
    B, N, h, w = class_stacks.shape
    ang = np.asarray(angles, np.float32).reshape(B, N)
    shf = np.asarray(shifts, np.float32).reshape(B, N, 2)
    out = {}
    for kind in sr_types:
        if kind == "aug":
            if max_stacks is None:
                x_class = sr.augmented_superresolution_batched(class_stacks, ang, shf)
                x_max = None
            else:
                # per image: class solve, then max solve, on one shared step counter -> interleave the stacks
                both = torch.stack([class_stacks, max_stacks], dim=1).reshape(2 * B, N, h, w).contiguous()
                x = sr.augmented_superresolution_batched(both, np.repeat(ang, 2, axis=0), np.repeat(shf, 2, axis=0))
                x = x.reshape(B, 2, *x.shape[1:])
                x_class, x_max = x[:, 0].contiguous(), x[:, 1].contiguous()
        elif kind in ("max", "mean"):
            x_class = sr.backproject_batched(class_stacks, ang, shf, kind)
            x_max = None if max_stacks is None else sr.backproject_batched(max_stacks, ang, shf, kind)
        else:
            raise ValueError("SR_type must be either 'aug', 'mean' or 'max'")
        out[kind] = _threshold_batched(x_class, class_id, th_factor, th_mask=x_max)
    return out


def run_files(sr: Superresolution, paths: Iterable[str], num_aug: int, class_id: int, th_factor: float, batch: int = 64,
              global_normalize: bool = True, sr_types: Sequence[str] = ("aug", "max", "mean"),
              on_skip: Optional[Callable[[str], None]] = None):
    """Generator over batches of hdf5 augmented-copies files (layout: augmentation_utils.py:117-136).
    Yields one BatchResult per `batch` valid files.  Files of one batch must share the OPM mode's
    structure (all with or all without max_masks), as a directory written by
    generate_augmented_copies.py does; a change of structure simply closes the current batch."""
    torch = _lib._torch()
    pending = []          # (path, class_masks, max_masks, angles, shifts, filename)
    skipped: List[str] = []

    def flush():
        nonlocal pending, skipped
        if not pending:
            return None
        cls = torch.stack([p[1].reshape(p[1].shape[0], p[1].shape[1], p[1].shape[2]) for p in pending]).contiguous()
        has_max = pending[0][2] is not None
        mx = torch.stack([p[2].reshape(p[2].shape[0], p[2].shape[1], p[2].shape[2]) for p in pending]).contiguous() if has_max else None
        ang = np.stack([np.asarray(p[3], np.float32) for p in pending])
        shf = np.stack([np.asarray(p[4], np.float32) for p in pending])
        masks = solve_stacks(sr, cls, mx, ang, shf, class_id, th_factor, sr_types)
        res = BatchResult(filenames=[p[5] for p in pending], paths=[p[0] for p in pending], skipped=skipped,
                          aug=masks.get("aug"), max=masks.get("max"), mean=masks.get("mean"))
        pending, skipped = [], []
        return res

    for path in paths:
        try:
            cm, mm, ang, shf, name = load_SR_data(path, num_aug=num_aug, global_normalize=global_normalize)
        except Exception:
            skipped.append(path)
            if on_skip is not None:
                on_skip(path)
            continue
        if pending and ((pending[0][2] is None) != (mm is None) or pending[0][1].shape != cm.shape):
            r = flush()
            if r is not None:
                yield r
        pending.append((path, cm, mm, ang, shf, name))
        if len(pending) >= batch:
            yield flush()
    r = flush()
    if r is not None:
        yield r
    elif skipped:
        yield BatchResult(filenames=[], paths=[], skipped=skipped)


def iou_table(true_masks, pred_masks, class_id):
    """The reference's per-image IoU pairs (SR_single_class.py:109-120) for a batch on the device:
    returns (iou_single [B], iou_with_bg [B]) float64 arrays."""
    return (compute_IoU_batched(true_masks, pred_masks, class_id, include_bg=False),
            compute_IoU_batched(true_masks, pred_masks, class_id, include_bg=True))
