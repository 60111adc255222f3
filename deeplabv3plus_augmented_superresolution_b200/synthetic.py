"""Synthetic augmented-copies generator.

There is no network for VOC or DeepLab weights, so benchmarks and tests use analytic stand-ins for
what generate_augmented_copies.py writes: for every image a random union of ellipses on the 512x512
canvas, and for every copy k the argmax-OPM-like hard mask the DeepLab forward would have produced
for the rotated/translated input, sampled at the centre of each low-resolution cell:

    y_k[i, j] = value * [ R_k (Z_ij - d_k) inside the shape ],   Z_ij = (4j + 1.5, 4i + 1.5)

Angles and shifts are drawn exactly as create_augmented_copies does (augmentation_utils.py:14-20):
legacy MT19937 stream, uniform(-angle_max, angle_max, N) then uniform(-shift_max, shift_max, (N, 2)),
element 0 forced to identity, cast to float32; the stream continues across images
(generate_augmented_copies.py:43,88-91).  torch is used only as array plumbing.
"""
from __future__ import annotations

import numpy as np
import torch


def draw_angles_shifts(rng: np.random.RandomState, num_aug: int, angle_max: float, shift_max: float):
    angles = rng.uniform(-angle_max, angle_max, num_aug)
    shifts = rng.uniform(-shift_max, shift_max, (num_aug, 2))
    angles[0] = 0
    shifts[0] = np.array([0, 0])
    return angles.astype("float32"), shifts.astype("float32")


def make_augmented_copies(num_images: int, num_aug: int = 100, feature_size=(128, 128), output_size=(512, 512),
                          angle_max: float = 0.15, shift_max: float = 80, seed: int = 1234, value: float = 1.0,
                          device="cpu", n_shapes: int = 2):
    """Returns (copies [B,N,h,w] f32 on `device`, angles [B,N] f32 ndarray, shifts [B,N,2] f32 ndarray)."""
    h, w = feature_size
    H, W = output_size
    sy, sx = H / h, W / w
    rng = np.random.RandomState(seed)          # the reference's np.random.seed(1234) stream
    shape_rng = np.random.RandomState(seed + 7919)
    angles = np.empty((num_images, num_aug), np.float32)
    shifts = np.empty((num_images, num_aug, 2), np.float32)
    for b in range(num_images):
        angles[b], shifts[b] = draw_angles_shifts(rng, num_aug, angle_max, shift_max)
    # ellipse parameters per image: centre, radii, orientation
    cx = shape_rng.uniform(0.3 * W, 0.7 * W, (num_images, n_shapes))
    cy = shape_rng.uniform(0.3 * H, 0.7 * H, (num_images, n_shapes))
    ra = shape_rng.uniform(0.08 * W, 0.28 * W, (num_images, n_shapes))
    rb = shape_rng.uniform(0.08 * H, 0.28 * H, (num_images, n_shapes))
    ph = shape_rng.uniform(0, np.pi, (num_images, n_shapes))

    dev = torch.device(device)
    jj = (torch.arange(w, device=dev, dtype=torch.float64) * sx + (sx - 1) / 2)[None, None, :]   # Z_x
    ii = (torch.arange(h, device=dev, dtype=torch.float64) * sy + (sy - 1) / 2)[None, :, None]   # Z_y
    out = torch.empty((num_images, num_aug, h, w), dtype=torch.float32, device=dev)
    for b in range(num_images):
        a = torch.as_tensor(angles[b], device=dev, dtype=torch.float64)[:, None, None]
        d = torch.as_tensor(shifts[b], device=dev, dtype=torch.float64)
        c, s = torch.cos(a), torch.sin(a)
        xoff = ((W - 1) - (c * (W - 1) - s * (H - 1))) / 2
        yoff = ((H - 1) - (s * (W - 1) + c * (H - 1))) / 2
        qx = jj - d[:, 0, None, None]
        qy = ii - d[:, 1, None, None]
        px = c * qx - s * qy + xoff            # source point in the un-augmented frame
        py = s * qx + c * qy + yoff
        inside_canvas = (px >= 0) & (px <= W - 1) & (py >= 0) & (py <= H - 1) & \
                        (qx >= 0) & (qx <= W - 1) & (qy >= 0) & (qy <= H - 1)
        m = torch.zeros_like(px, dtype=torch.bool)
        for e in range(n_shapes):
            ux = px - cx[b, e]
            uy = py - cy[b, e]
            cp, sp = np.cos(ph[b, e]), np.sin(ph[b, e])
            ex = (cp * ux + sp * uy) / ra[b, e]
            ey = (-sp * ux + cp * uy) / rb[b, e]
            m |= (ex * ex + ey * ey) <= 1.0
        out[b] = (m & inside_canvas).to(torch.float32) * value
    return out, angles, shifts


# ---------------------------------------------------------------------------------------------------------------
# stand-ins for test_SR.py's inputs: an image with a ground-truth mask and an upstream "model"
# ---------------------------------------------------------------------------------------------------------------
def make_test_image(size=(512, 512), class_id: int = 8, seed: int = 5):
    """An RGB image in [0,1] whose red channel marks a union of two ellipses, and its label image (values {0, class_id}).
    Stands in for test_images/test_cat.jpg + test_cat_gt.png (there are no DeepLab weights here to segment a real cat)."""
    H, W = size
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    m = np.zeros((H, W), bool)
    for _ in range(2):
        cx, cy = rng.uniform(0.35 * W, 0.65 * W), rng.uniform(0.35 * H, 0.65 * H)
        ra, rb, ph = rng.uniform(0.12 * W, 0.25 * W), rng.uniform(0.12 * H, 0.25 * H), rng.uniform(0, np.pi)
        ux, uy = xx - cx, yy - cy
        ex, ey = (np.cos(ph) * ux + np.sin(ph) * uy) / ra, (-np.sin(ph) * ux + np.cos(ph) * uy) / rb
        m |= (ex * ex + ey * ey) <= 1.0
    img = np.empty((H, W, 3), np.float32)
    img[..., 0] = np.where(m, 0.9, 0.1)
    img[..., 1] = 0.3 + 0.2 * rng.rand(H, W)
    img[..., 2] = np.where(m, 0.2, 0.6)
    return img, (m.astype(np.int32) * class_id)[..., None]


class SyntheticSegmenter:
    """The upstream producer interface of compute_augmented_feature_maps (`model.predict(images, batch_size) ->
    [n, H/4, W/4, K] logits`, model.py in the reference, outside this repo).  Class `class_id` wins where the 4x4-pooled
    red channel exceeds one half, background elsewhere; a fixed pseudo-random ripple keeps the other channels distinct."""

    def __init__(self, classes: int = 21, class_id: int = 8, stride: int = 4):
        self.classes, self.class_id, self.stride = classes, class_id, stride

    def predict(self, images, batch_size=16):
        x = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.asarray(images, np.float32))
        x = x.to(torch.float32)
        n, H, W, _ = x.shape
        s = self.stride
        red = x[..., 0].reshape(n, H // s, s, W // s, s).mean(dim=(2, 4))
        k = torch.arange(self.classes, device=x.device, dtype=torch.float32)
        logits = 0.05 * torch.sin(k[None, None, None, :] * 1.7 + red[..., None] * 3.0)
        logits[..., 0] += 2.0 * (0.5 - red)
        logits[..., self.class_id] += 2.0 * (red - 0.5)
        return logits.contiguous()
