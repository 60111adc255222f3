"""Second, independent restatement of the operator semantics in vectorised
torch-CPU ops.  TEST INFRASTRUCTURE: it exists only to cross-check oracle/asr_oracle.c, which was
written as scalar C loops, against a differently-structured implementation of the same definitions
(agreement to fp32 rounding, not bit-exactness).  Not imported by the product.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def rotate_matrix(angle: float, H: int, W: int) -> torch.Tensor:
    c, s = math.cos(angle), math.sin(angle)
    xo = ((W - 1) - (c * (W - 1) - s * (H - 1))) / 2.0
    yo = ((H - 1) - (s * (W - 1) + c * (H - 1))) / 2.0
    return torch.tensor([c, -s, xo, s, c, yo, 0.0, 0.0], dtype=torch.float64)


def translate_matrix(dx: float, dy: float) -> torch.Tensor:
    return torch.tensor([1.0, 0.0, -dx, 0.0, 1.0, -dy, 0.0, 0.0], dtype=torch.float64)


def invert(t: torch.Tensor) -> torch.Tensor:
    m = torch.cat([t, torch.ones(1, dtype=t.dtype)]).reshape(3, 3)
    inv = torch.linalg.inv(m)
    inv = inv / inv[2, 2]
    return inv.reshape(-1)[:8]


def transform(img: torch.Tensor, t: torch.Tensor, nearest: bool = False) -> torch.Tensor:
    """ImageProjectiveTransformV3 on one [H,W] image, fill 0 (A.1)."""
    H, W = img.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    proj = t[6] * xs + t[7] * ys + 1.0
    ix = (t[0] * xs + t[1] * ys + t[2]) / proj
    iy = (t[3] * xs + t[4] * ys + t[5]) / proj

    def rd(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = img[yy.clamp(0, H - 1).long(), xx.clamp(0, W - 1).long()]
        return torch.where(ok, v, torch.zeros_like(v))

    if nearest:
        rnd = lambda v: torch.sign(v) * torch.floor(v.abs() + 0.5)   # std::round: half away from zero
        return rd(rnd(iy), rnd(ix))
    x0, y0 = torch.floor(ix), torch.floor(iy)
    wx1, wy1 = (ix - x0).to(img.dtype), (iy - y0).to(img.dtype)
    top = (1 - wx1) * rd(y0, x0) + wx1 * rd(y0, x0 + 1)
    bot = (1 - wx1) * rd(y0 + 1, x0) + wx1 * rd(y0 + 1, x0 + 1)
    return (1 - wy1) * top + wy1 * bot


def resize(img: torch.Tensor, size) -> torch.Tensor:
    """tf.image.resize bilinear, half-pixel centres, no antialias == F.interpolate(align_corners=False)."""
    return F.interpolate(img[None, None], size=tuple(size), mode="bilinear", align_corners=False, antialias=False)[0, 0]


def loss_and_grad(x, copies, angles, shifts, lambda_df, lambda_tv, lambda_l2, lambda_l1):
    """A.3 forward + A.4 gradient (TensorFlow's registered warp gradient, not the exact adjoint)."""
    H, W = x.shape
    N, h, w = copies.shape
    x = x.double(); copies = copies.double()
    g = torch.zeros_like(x)
    df = 0.0
    for k in range(N):
        R = rotate_matrix(float(angles[k]), H, W)
        T = translate_matrix(float(shifts[k, 0]), float(shifts[k, 1]))
        z = transform(transform(x, R), T)
        D = resize(z, (h, w))
        r = D - copies[k]
        df = df + (r * r).sum()
        # ResizeBilinearGrad = exact transpose of the resize
        glr = (2.0 * lambda_df * r).clone().requires_grad_(False)
        zz = torch.zeros((H, W), dtype=torch.float64, requires_grad=True)
        (resize(zz, (h, w)) * glr).sum().backward()
        ghr = zz.grad
        g = g + transform(transform(ghr, invert(T)), invert(R))
    dy = torch.zeros_like(x); dy[:-1] = x[1:] - x[:-1]
    dx = torch.zeros_like(x); dx[:, :-1] = x[:, 1:] - x[:, :-1]
    tv = dy.abs().sum() + dx.abs().sum()
    sy, sx = torch.sign(dy), torch.sign(dx)
    gtv = -sy - sx
    gtv[1:] += sy[:-1]
    gtv[:, 1:] += sx[:, :-1]
    g = g + lambda_tv * gtv + 2.0 * lambda_l2 * x
    loss = lambda_df * df + lambda_tv * tv + lambda_l2 * (x * x).sum()
    if lambda_l1 > 0:
        g = g + lambda_l1 * torch.sign(x)
        loss = loss + lambda_l1 * x.abs().sum()
    return float(loss), g
