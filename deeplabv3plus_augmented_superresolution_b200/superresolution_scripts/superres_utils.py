"""Drop-in for superresolution_scripts/superres_utils.py (reference :1-273) on libasr.

Same function names, arguments, return types and error behaviour.  Device work (global min-max
normalisation, thresholding, the solves) goes through the C ABI; path helpers are plain Python as in
the reference.  Reference quirks that would crash are kept harmless: a
`max_masks` of None is treated as empty.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .. import _lib, hdf5_lite
from .superresolution import Superresolution, _as_device_stack

try:  # the reference imports h5py; use it when present, else the bundled spec-level implementation
    import h5py as _h5
except Exception:  # pragma: no cover - h5py is absent in the build image
    _h5 = hdf5_lite


def get_img_paths(image_list_path, image_folder, is_png=False, sort=True):
    """reference :9-29"""
    ext = ".jpg" if not is_png else ".png"
    paths = [os.path.join(image_folder, line.rstrip() + ext)
             for line in open(image_list_path)]
    if sort:
        paths = sorted(paths, key=lambda p: int(
            os.path.basename(p).split('.')[0]))
    return paths


def class_in_image(image_path, class_id, image_size=(512, 512)):
    """reference :32-38"""
    from ..utils import load_image
    mask_path = image_path.replace(
        "JPEGImages", "SegmentationClassAug").replace("jpg", "png")
    mask = load_image(mask_path, image_size=image_size,
                      normalize=False, is_png=True, resize_method="nearest")
    return np.any(mask == class_id)


def filter_images_by_class(path_list, filter_class_id, num_images=None, image_size=(512, 512)):
    """reference :41-53"""
    max_images = num_images if num_images is not None else len(path_list)
    image_paths = []
    for path in path_list:
        if len(image_paths) == max_images:
            break
        if class_in_image(path, class_id=filter_class_id, image_size=image_size):
            image_paths.append(path)
    return image_paths


def min_max_normalization(image, new_min=0.0, new_max=255.0, global_min=None, global_max=None):
    """reference :56-62 (host NumPy, same expression order; fp32 in, fp32 out)."""
    image = np.asarray(image)
    mn = image.min() if global_min is None else global_min
    mx = image.max() if global_max is None else global_max
    num = (image - mn) * (new_max - new_min)
    den = (mx - mn) if (mx - mn) != 0 else 1.0
    return new_min + (num / den)


def list_precomputed_data_paths(root_dir, sort=False):
    """reference :93-105"""
    paths = []
    for path, subdirs, files in os.walk(root_dir):
        for filename in files:
            if filename.endswith(".hdf5"):
                paths.append(os.path.join(path, filename))
    if sort:
        paths = sorted(paths, key=lambda p: int(
            os.path.basename(p).split('.')[0]))   # int() accepts the "2007_000032" underscore form
    return paths


def check_hdf5_validity(file, num_aug=100):
    """reference :108-115"""
    for keys in file:
        num = file[keys].shape[0]
        if num < num_aug:
            return False
    return True


def _normalize_stack_device(stack):
    """Global min-max normalisation of a whole [N,h,w] stack to [0,1] on the device
    (reference :186-194: tf.reduce_min/max over all copies, then min_max_normalization per image)."""
    torch = _lib._torch()
    L = _lib.lib()
    out = torch.empty_like(stack)
    ws = torch.empty(2, dtype=torch.float32, device=stack.device)
    with torch.cuda.device(stack.device):
        _lib.check(L.asr_minmax_normalize(stack.data_ptr(), stack.numel(), 0.0, 1.0, out.data_ptr(), ws.data_ptr(),
                                          _lib._stream_ptr(torch)))
    return out


def threshold_image(image, th_value, th_factor=.15, th_mask=None):
    """reference :118-139 -> int32 ndarray of `image`'s shape with values {0, th_value}."""
    torch = _lib._torch()
    L = _lib.lib()
    shape = tuple(np.shape(image)) if not isinstance(image, torch.Tensor) else tuple(image.shape)
    x = (image if isinstance(image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)))
    x = x.to(device="cuda", dtype=torch.float32).contiguous()
    m = None
    if th_mask is not None:
        m = (th_mask if isinstance(th_mask, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(th_mask, dtype=np.float32)))
        m = m.to(device="cuda", dtype=torch.float32).contiguous()
    out = torch.empty(x.numel(), dtype=torch.int32, device=x.device)
    ws = torch.empty(2, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.asr_threshold(x.data_ptr(), 1, x.numel(), int(th_value), float(th_factor),
                                   None if m is None else m.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                   _lib._stream_ptr(torch)))
    return out.cpu().numpy().reshape(shape)


def normalize_coefficients(coeff_dict):
    """reference :142-151"""
    normalizer = np.sum(list(coeff_dict.values()))
    return {key: (value / normalizer) for (key, value) in coeff_dict.items()}


def load_SR_data(filepath, num_aug=100, global_normalize=True):
    """reference :154-210.  Returns (class_masks, max_masks | None, angles, shifts, filename); the two
    stacks come back as CUDA float32 tensors [N,h,w,1] (where the reference returns tf.Tensors)."""
    torch = _lib._torch()
    file = _h5.File(f"{filepath}", "r")

    if not check_hdf5_validity(file, num_aug=num_aug):
        file.close()
        raise Exception(f"File: {filepath} is invalid")

    filename = file.attrs["filename"]
    mode = file.attrs["mode"]
    angles = file["angles"][:num_aug]
    shifts = file["shifts"][:num_aug]

    class_masks = torch.from_numpy(np.ascontiguousarray(file["class_masks"][:num_aug], dtype=np.float32)).cuda()
    max_masks = None

    if mode != "slice":
        if global_normalize:
            class_masks = _normalize_stack_device(class_masks)
        else:  # per-image min/max (global_min=None path of min_max_normalization)
            class_masks = torch.stack([_normalize_stack_device(c.contiguous()) for c in class_masks])

    if mode == "slice_max":
        max_masks = torch.from_numpy(np.ascontiguousarray(file["max_masks"][:num_aug], dtype=np.float32)).cuda()
        if global_normalize:
            max_masks = _normalize_stack_device(max_masks)
        else:
            max_masks = torch.stack([_normalize_stack_device(c.contiguous()) for c in max_masks])

    file.close()
    return class_masks, max_masks, angles, shifts, filename


def _save_img(path, array, scale=True):
    """tf.keras.utils.save_img(path, x, scale=True) for a single-channel image (array_to_img rescales
    x - min to [0,255] by its max)."""
    from PIL import Image
    a = np.asarray(array, dtype=np.float32)
    if a.ndim == 3:
        a = a[..., 0]
    if scale:
        a = a - a.min()
        mx = a.max()
        if mx != 0:
            a = a / mx
        a = a * 255.0
    Image.fromarray(a.astype(np.uint8), mode="L").save(path)


def compute_SR(superresolution_obj: Superresolution, class_masks, angles, shifts, filename, dest_folder,
               SR_type="aug", max_masks=[], save_intermediate_output=False, save_final_output=False, class_id=8,
               th_factor=0.15):
    """reference :213-273 -> thresholded [H,W,1] int32 mask with values {0, class_id}."""
    if SR_type not in ["aug", "mean", "max"]:   # the reference's assert is a tuple and never fires (:235-236)
        raise ValueError("SR_type must be either 'aug', 'mean' or 'max'")

    out_folder = os.path.join(dest_folder, f"{SR_type}_SR")
    if not os.path.exists(out_folder):
        os.makedirs(out_folder)

    if SR_type == "aug":
        SR_function = superresolution_obj.augmented_superresolution
    elif SR_type == "mean":
        SR_function = superresolution_obj.mean_superresolution
    elif SR_type == "max":
        SR_function = superresolution_obj.max_superresolution

    n_max = 0 if max_masks is None else len(max_masks)       # reference :253 crashes on None (Appendix B-2)
    target_image_max = None
    if n_max == len(class_masks):
        if SR_type == "aug" and not superresolution_obj.verbose:
            # the class solve and the max solve of this image as one two-image call (same step offsets, same results)
            from ..batch_runner import solve_class_and_max
            target_image_class, target_image_max = solve_class_and_max(superresolution_obj, class_masks, max_masks, angles, shifts)
        else:
            target_image_class, _ = SR_function(class_masks, angles, shifts)
            target_image_max, _ = SR_function(max_masks, angles, shifts)
        th_mask = threshold_image(
            target_image_class, class_id, th_mask=target_image_max)
    else:
        target_image_class, _ = SR_function(class_masks, angles, shifts)
        th_mask = threshold_image(
            target_image_class, class_id, th_factor=th_factor)

    if save_intermediate_output:
        _save_img(f"{out_folder}/{filename}_class.png", target_image_class, scale=True)
        if target_image_max is not None:
            _save_img(f"{out_folder}/{filename}_max.png", target_image_max, scale=True)

    if save_final_output:
        _save_img(f"{out_folder}/{filename}_{SR_type}_SR.png", th_mask, scale=True)

    return th_mask
