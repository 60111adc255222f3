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
