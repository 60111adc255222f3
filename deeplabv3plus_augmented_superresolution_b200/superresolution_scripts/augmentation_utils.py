"""Drop-in for superresolution_scripts/augmentation_utils.py (reference :1-138) on libasr.

create_augmented_copies draws angles/shifts from the global NumPy RNG exactly as the reference does
(:14-20) and runs the rotate->translate warp on the device (asr_warp_affine); the OPM extraction of
compute_augmented_feature_maps (:80-115) is asr_opm_extract.  `model` is the upstream producer
(DeepLabV3+ in the reference, outside this repo): any object with predict(images, batch_size=...)
returning [N,h,w,K] logits as a NumPy array or a torch tensor.
"""
from __future__ import annotations

import ctypes as C
import gc
import os

import numpy as np

from .. import _lib, hdf5_lite
from ..utils import load_image

try:
    import h5py as _h5
except Exception:  # pragma: no cover
    _h5 = hdf5_lite


def _draw(num_aug, angle_max, shift_max):
    angles = np.random.uniform(-angle_max, angle_max, num_aug)
    shifts = np.random.uniform(-shift_max, shift_max, (num_aug, 2))
    # First sample is not augmented
    angles[0] = 0
    shifts[0] = np.array([0, 0])
    return angles.astype("float32"), shifts.astype("float32")


def warp_copies(image, angles, shifts, interpolation="bilinear"):
    """tfa.image.rotate then tfa.image.translate of one [H,W,C] image for every (angle, shift):
    CUDA tensor [N,H,W,C] (reference :22-25; also check_robustness.py:44-50 with 'nearest')."""
    torch = _lib._torch()
    L = _lib.lib()
    img = image if isinstance(image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32))
    img = img.to(device="cuda", dtype=torch.float32).contiguous()
    if img.dim() == 2:
        img = img[..., None].contiguous()
    H, W, Cc = img.shape
    ang = np.ascontiguousarray(np.asarray(angles, dtype=np.float32).reshape(-1))
    shf = np.ascontiguousarray(np.asarray(shifts, dtype=np.float32).reshape(-1, 2))
    n = ang.shape[0]
    out = torch.empty((n, H, W, Cc), dtype=torch.float32, device=img.device)
    need = C.c_size_t()
    _lib.check(L.asr_warp_affine_workspace_bytes(n, H, W, Cc, C.byref(need)))
    ws = _lib._aux_ws.get(need.value, img.device)
    with torch.cuda.device(img.device):
        _lib.check(L.asr_warp_affine_ws(img.data_ptr(), ang.ctypes.data_as(_lib._fp), shf.ctypes.data_as(_lib._fp), n, H, W, Cc,
                                        _lib.INTERP[interpolation.lower()], out.data_ptr(), ws.data_ptr(), ws.numel(),
                                        _lib._stream_ptr(torch)))
    return out


def create_augmented_copies(image, num_aug, angle_max, shift_max):
    """reference :11-27 -> (copies CUDA [num_aug,H,W,C], angles f32 [num_aug], shifts f32 [num_aug,2])."""
    angles, shifts = _draw(num_aug, angle_max, shift_max)
    return warp_copies(image, angles, shifts, "bilinear"), angles, shifts


def create_augmented_copies_chunked(image, num_aug, angle_max, shift_max, chunk_size=100):
    """reference :30-59: same draws, warp in chunks, NumPy result."""
    if (num_aug % chunk_size) != 0:
        raise Exception("Num aug must be a multiple of 50")
    num_chunks = num_aug // chunk_size
    angles, shifts = _draw(num_aug, angle_max, shift_max)
    chunks = [warp_copies(image, a, s, "bilinear").cpu().numpy()
              for a, s in zip(np.split(angles, num_chunks), np.split(shifts, num_chunks))]
    return np.concatenate(chunks, axis=0), angles, shifts


def extract_opm(predictions, filter_class_id, mode):
    """OPM extraction of a whole prediction stack on the device (reference :80-115).
    predictions [N,h,w,K] -> (class_masks CUDA [N,h,w,1], max_masks CUDA [N,h,w,1] | None)."""
    torch = _lib._torch()
    L = _lib.lib()
    p = predictions if isinstance(predictions, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(predictions, dtype=np.float32))
    p = p.to(device="cuda", dtype=torch.float32).contiguous()
    n, h, w, K = p.shape
    m = "argmax" if mode not in ("slice", "slice_max") else mode     # the reference's final `else` branch
    cls = torch.empty((n, h, w, 1), dtype=torch.float32, device=p.device)
    mx = torch.empty((n, h, w, 1), dtype=torch.float32, device=p.device) if m == "slice_max" else None
    ws = torch.empty(2 * n, dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(L.asr_opm_extract(p.data_ptr(), n, h, w, K, int(filter_class_id), _lib.OPM_MODES[m], cls.data_ptr(),
                                     None if mx is None else mx.data_ptr(), ws.data_ptr(), _lib._stream_ptr(torch)))
    return cls, mx


def compute_augmented_feature_maps(image_path, model, filter_class_id, mode="slice", num_aug=100,
                                   angle_max=0.5, shift_max=30, image_size=(512, 512), batch_size=16, dest_folder=None):
    """reference :62-138 -> (class_masks list of [h,w,1] f32 arrays, max_masks list, angles, shifts, image_name)."""
    torch = _lib._torch()
    image_name = os.path.splitext(os.path.basename(image_path))[0]

    image = load_image(image_path, image_size=image_size, normalize=True)
    augmented_copies, angles, shifts = create_augmented_copies(image, num_aug=num_aug, angle_max=angle_max,
                                                               shift_max=shift_max)

    predictions = model.predict(augmented_copies, batch_size=batch_size)
    _ = gc.collect()

    cls, mx = extract_opm(predictions, filter_class_id, mode)
    cls_np = cls.cpu().numpy()
    class_masks = [cls_np[i] for i in range(cls_np.shape[0])]
    max_masks = []
    if mx is not None:
        mx_np = mx.cpu().numpy()
        max_masks = [mx_np[i] for i in range(mx_np.shape[0])]

    if dest_folder is not None:
        if not os.path.exists(dest_folder):
            os.makedirs(dest_folder)

        file = _h5.File(f"{dest_folder}/{image_name}.hdf5", "w")
        file.create_dataset("class_masks", data=class_masks)
        if mode == "slice_max":
            file.create_dataset("max_masks", data=max_masks)
        file.create_dataset("angles", data=angles)
        file.create_dataset("shifts", data=shifts)
        file.attrs["filename"] = image_name
        file.attrs["mode"] = mode
        file.attrs["angle_max"] = angle_max
        file.attrs["shift_max"] = shift_max
        file.close()

    return class_masks, max_masks, angles, shifts, image_name
