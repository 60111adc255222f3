"""Host helpers mirroring the parts of the reference's utils.py the hot path touches
(load_image :94-112, create_mask :115-119, single_class_IOU :180-204, Mean_IOU :151-177,
compute_IoU :207-230).  Plotting and the Keras training metrics are out of scope."""
from __future__ import annotations

import numpy as np


def _resize(img: np.ndarray, size, method: str) -> np.ndarray:
    """tf.image.resize(img, size, method) for method in {bilinear, nearest}: half-pixel centres,
    no antialiasing, fp32."""
    img = np.asarray(img, dtype=np.float32)
    H, W = size
    h, w = img.shape[:2]
    if method == "nearest":
        # ResizeNearestNeighbor(half_pixel_centers=True): src = min(floor((o + 0.5) * scale), in - 1)
        ys = np.minimum(np.floor((np.arange(H, dtype=np.float32) + np.float32(0.5)) * np.float32(h / H)).astype(np.int64), h - 1)
        xs = np.minimum(np.floor((np.arange(W, dtype=np.float32) + np.float32(0.5)) * np.float32(w / W)).astype(np.int64), w - 1)
        return img[ys][:, xs]

    def weights(out, inn):
        scale = np.float32(inn) / np.float32(out)
        src = (np.arange(out, dtype=np.float32) + np.float32(0.5)) * scale - np.float32(0.5)
        f = np.floor(src)
        lo = np.maximum(f.astype(np.int64), 0)
        hi = np.minimum(np.ceil(src).astype(np.int64), inn - 1)
        return lo, hi, (src - f).astype(np.float32)

    ylo, yhi, yl = weights(H, h)
    xlo, xhi, xl = weights(W, w)
    xl = xl.reshape((1, W) + (1,) * (img.ndim - 2))
    yl = yl.reshape((H, 1) + (1,) * (img.ndim - 2))
    top = img[ylo][:, xlo] + (img[ylo][:, xhi] - img[ylo][:, xlo]) * xl
    bot = img[yhi][:, xlo] + (img[yhi][:, xhi] - img[yhi][:, xlo]) * xl
    return (top + (bot - top) * yl).astype(np.float32)


def load_image(img_path, image_size=None, normalize=True, is_png=False, resize_method="bilinear"):
    """reference utils.py:94-112 -> float32 [H,W,3] (jpeg) or [H,W,1] (png)."""
    from PIL import Image
    im = Image.open(img_path)
    if not is_png:
        image = np.asarray(im.convert("RGB"), dtype=np.uint8)
    else:
        # tf.image.decode_png(channels=1): palette PNGs yield the palette *indices*
        image = np.asarray(im if im.mode in ("P", "L") else im.convert("L"), dtype=np.uint8)[..., None]
    image = image.astype(np.float32)
    if image_size is not None:
        image = _resize(image, image_size, resize_method)
    if normalize:
        image = image / np.float32(255.0)
    return image.astype(np.float32)


def create_mask(pred_mask):
    """reference utils.py:115-119: argmax over the class axis, keep a trailing unit axis (ties -> lowest index)."""
    pred_mask = np.argmax(np.asarray(pred_mask), axis=-1)
    return pred_mask[..., np.newaxis]


def Mean_IOU(y_true, y_pred):
    """reference utils.py:151-177"""
    t = np.squeeze(np.asarray(y_true)).astype(np.int32)
    p = np.squeeze(np.asarray(y_pred)).astype(np.int32)
    labels = [l for l in np.unique(t) if l != 255]
    ious = []
    for i in labels:
        tl, pl = t == i, p == i
        ious.append(np.sum(tl & pl) / np.sum(tl | pl))
    return np.float64(np.mean(ious))


def single_class_IOU(y_true, y_pred, class_id, include_bg):
    """reference utils.py:180-204: NaN (empty union) entries are dropped before the mean."""
    t = np.squeeze(np.asarray(y_true))
    p = np.squeeze(np.asarray(y_pred))
    classes = [class_id]
    if include_bg:
        classes.append(0)
        t = np.where(t != class_id, 0, t)
    ious = []
    for i in classes:
        tl, pl = t == i, p == i
        union = np.sum(tl | pl)
        ious.append(np.sum(tl & pl) / union if union else np.nan)
    ious = np.array(ious, dtype=np.float64)
    return np.float64(np.mean(ious[~np.isnan(ious)])) if np.any(~np.isnan(ious)) else np.float64(np.nan)


def compute_IoU(true_image, image, img_size=(512, 512), class_id=None, include_bg=False):
    """reference utils.py:207-230"""
    true_image = np.reshape(np.asarray(true_image), (img_size[0] * img_size[1], 1))
    image = np.reshape(np.asarray(image), (img_size[0] * img_size[1], 1))
    if class_id is not None:
        return single_class_IOU(true_image, image, class_id, include_bg)
    return Mean_IOU(true_image, image)


def compute_IoU_batched(true_images, images, class_id, include_bg=False):
    """compute_IoU (single-class branch) for B label images at once on the device.
    true_images / images: CUDA int32 tensors [B,...] with the same number of pixels per image.
    Returns a float64 ndarray [B]; NaN where every union is empty, as the reference's mean of nothing."""
    from . import _lib
    torch = _lib._torch()
    L = _lib.lib()
    t = true_images.to(device="cuda", dtype=torch.int32).contiguous()
    p = images.to(device="cuda", dtype=torch.int32).contiguous()
    B = t.shape[0]
    n = t[0].numel()
    assert p.shape[0] == B and p[0].numel() == n
    counts = torch.empty((B, 4), dtype=torch.int64, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(L.asr_iou_counts(t.data_ptr(), p.data_ptr(), B, n, int(class_id), int(bool(include_bg)), counts.data_ptr(),
                                    _lib._stream_ptr(torch)))
    c = counts.cpu().numpy().astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou_c = np.where(c[:, 1] > 0, c[:, 0] / c[:, 1], np.nan)
        iou_b = np.where(c[:, 3] > 0, c[:, 2] / c[:, 3], np.nan)
    if not include_bg:
        return iou_c
    stack = np.stack([iou_c, iou_b], 1)
    out = np.full(B, np.nan)
    ok = ~np.isnan(stack)
    has = ok.any(1)
    out[has] = np.nansum(stack[has], 1) / ok[has].sum(1)
    return out
