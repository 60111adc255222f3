"""Third-party pins of the oracle's operator semantics (CPU).

The reference's arithmetic lives in tensorflow 2.7 / tensorflow-addons 0.15, which cannot be installed here or on the GPU
box (profiles/r02_tf_install_attempt.log), so the oracle stays "parity unpinned" with respect to TensorFlow itself.  What
CAN be checked without TensorFlow is that the oracle's building blocks agree with independent, widely used
implementations of the same published operators -- none of them written by this repo:

  * ImageProjectiveTransformV3 (bilinear, CONSTANT fill 0)  ==  scipy.ndimage.affine_transform(order=1, mode="grid-constant")
  * tf.image.resize(bilinear, half-pixel centres, no antialias)  ==  torch.nn.functional.interpolate(bilinear, align_corners=False)
  * tf.image.resize(nearest, half-pixel centres)                 ==  torch interpolate(mode="nearest-exact")
  * tf.linalg.inv of the 3x3 transform                            ==  numpy.linalg.inv (float64)
  * Keras Adam / AMSGrad                                          ==  torch.optim.Adam up to the placement of epsilon
  * tf.image.image_gradients TV + L2 gradients                    ==  torch.autograd on the same expression

The one piece with no third-party counterpart is TensorFlow's registered gradient of the warp (the forward op with the
inverted transform instead of the adjoint); tests/test_oracle.py::test_gradient_is_tensorflow_not_adjoint
documents its distance from the true adjoint.
"""
import numpy as np
import pytest
import torch
from scipy import ndimage


def _scipy_transform(img, t):
    """out[y, x] = bilinear(img, (t3 x + t4 y + t5, t0 x + t1 y + t2)) with every tap outside the image read as 0."""
    m = np.array([[t[4], t[3]], [t[1], t[0]]], np.float64)
    return ndimage.affine_transform(img.astype(np.float64), m, offset=[t[5], t[2]], order=1, mode="grid-constant", cval=0.0, prefilter=False)


@pytest.mark.parametrize("angle,shift", [(0.0, (0.0, 0.0)), (0.15, (7.3, -4.6)), (-0.15, (79.9, 80.0)), (1.3, (-20.5, 30.25)), (3.1, (0.5, 0.5))])
def test_rotate_then_translate_vs_scipy(oracle, angle, shift):
    rng = np.random.RandomState(0)
    img = rng.rand(1, 96, 128, 1).astype(np.float32)
    rot = oracle.rotate(img, [angle])
    out = oracle.translate(rot, [shift])[0, :, :, 0]
    tr = oracle.rotate_matrix(angle, 96, 128)
    tt = oracle.translate_matrix(*shift)
    ref = _scipy_transform(_scipy_transform(img[0, :, :, 0], tr), tt)
    # fp32 coordinates (|coordinate| ~ 100, ulp 8e-6) times unit-range gradients: 2e-5 is the fp32/fp64 evaluation gap
    np.testing.assert_allclose(out, ref, atol=3e-5)
    assert np.abs(out).max() > 0.1 or abs(shift[0]) > 70


def test_projective_fill_is_per_tap_zero(oracle):
    """Half a pixel outside the canvas the op returns half the edge pixel (each tap is zero-filled on its own), as scipy's
    grid-constant mode does and its plain constant mode does not."""
    img = np.ones((1, 8, 8, 1), np.float32)
    out = oracle.translate(img, [[0.5, 0.0]])[0, :, :, 0]
    np.testing.assert_allclose(out[:, 0], 0.5)
    np.testing.assert_allclose(out[:, 1:], 1.0)
    np.testing.assert_allclose(out, _scipy_transform(img[0, :, :, 0], oracle.translate_matrix(0.5, 0.0)), atol=1e-7)


@pytest.mark.parametrize("src,dst", [((128, 128), (512, 512)), ((512, 512), (128, 128)), ((24, 40), (96, 160)), ((96, 160), (24, 40))])
def test_resize_bilinear_vs_torch_interpolate(oracle, src, dst):
    rng = np.random.RandomState(1)
    a = rng.rand(1, src[0], src[1], 1).astype(np.float32)
    got = oracle.resize_bilinear(a, dst)[0, :, :, 0]
    ref = torch.nn.functional.interpolate(torch.from_numpy(a[0, :, :, 0])[None, None].double(), size=dst, mode="bilinear",
                                          align_corners=False, antialias=False)[0, 0].numpy()
    np.testing.assert_allclose(got, ref, atol=1e-6)


def test_nearest_resize_vs_torch_interpolate():
    """utils.load_image(resize_method='nearest') = tf.image.resize nearest with half-pixel centres."""
    from deeplabv3plus_augmented_superresolution_b200 import utils
    rng = np.random.RandomState(2)
    a = rng.randint(0, 21, (375, 500)).astype(np.float32)
    got = utils._resize(a[..., None], (512, 512), "nearest")[..., 0]
    ref = torch.nn.functional.interpolate(torch.from_numpy(a)[None, None], size=(512, 512), mode="nearest-exact")[0, 0].numpy()
    assert (got != ref).mean() < 1e-3        # identical except where fp32 (o+0.5)*scale lands exactly on an integer
    got2 = utils._resize(a[..., None], (128, 96), "bilinear")[..., 0]
    ref2 = torch.nn.functional.interpolate(torch.from_numpy(a)[None, None].double(), size=(128, 96), mode="bilinear", align_corners=False)[0, 0].numpy()
    np.testing.assert_allclose(got2, ref2, atol=1e-3)     # labels up to 20, fp32 source coordinates


def test_transform_inverse_vs_numpy(oracle):
    for angle in (0.0, 0.05, -0.15, 0.7, 3.0):
        t = oracle.rotate_matrix(angle, 512, 512)
        m = np.array([[t[0], t[1], t[2]], [t[3], t[4], t[5]], [0, 0, 1]], np.float64)
        inv = np.linalg.inv(m)
        inv = inv / inv[2, 2]
        np.testing.assert_allclose(oracle.invert_transform(t), inv.reshape(-1)[:8], atol=3e-4, rtol=1e-5)   # fp32 LU vs fp64; offsets are ~1e2


def test_adam_amsgrad_vs_torch_optim(oracle):
    """A constant-free check of the optimizer arithmetic: run the oracle with lambda_df = 0 (no data term), so the gradient is
    exactly 2*lambda_L2*x + lambda_tv*dTV, reproduce that gradient with torch.autograd and step torch.optim.Adam(amsgrad=True).
    Keras adds epsilon to sqrt(v_hat) before the bias correction, torch after it: for |g| >> 1e-7 the iterates agree to ~1e-6."""
    rng = np.random.RandomState(3)
    h = w = 8
    c = rng.rand(1, h, w).astype(np.float32)             # one (identity) copy; x0 = its bilinear upsample
    P = oracle.SolveParams(lambda_df=0.0, lambda_tv=0.0, lambda_l2=0.7, num_iter=12, lr_scheduler=False, learning_rate=1e-3, amsgrad=True)
    x, _ = oracle.augmented_superresolution(c, np.zeros(1, np.float32), np.zeros((1, 2), np.float32), P, output_size=(4 * h, 4 * w))
    x0 = oracle.resize_bilinear(c[:, :, :, None], (4 * h, 4 * w))[0, :, :, 0]
    xt = torch.from_numpy(x0.astype(np.float64)).requires_grad_(True)
    opt = torch.optim.Adam([xt], lr=1e-3, betas=(0.9, 0.999), eps=1e-7, amsgrad=True)
    for _ in range(12):
        opt.zero_grad()
        (0.7 * (xt ** 2).sum()).backward()
        opt.step()
    np.testing.assert_allclose(x[..., 0], xt.detach().numpy(), atol=2e-6)
    assert np.abs(x[..., 0] - x0).max() > 5e-3           # the twelve steps moved x by ~12 * lr


def test_tv_and_l2_gradient_vs_autograd(oracle):
    rng = np.random.RandomState(4)
    h = w = 8
    c = np.zeros((1, h, w), np.float32)
    x = rng.rand(4 * h, 4 * w).astype(np.float32)
    P = oracle.SolveParams(lambda_df=0.0, lambda_tv=0.3, lambda_l2=0.7, lambda_l1=0.05)
    loss, g = oracle.loss_and_grad(x, c, np.zeros(1, np.float32), np.zeros((1, 2), np.float32), P)
    xt = torch.from_numpy(x.astype(np.float64)).requires_grad_(True)
    tv = (xt[1:] - xt[:-1]).abs().sum() + (xt[:, 1:] - xt[:, :-1]).abs().sum()     # tf.image.image_gradients: forward differences
    lt = 0.3 * tv + 0.7 * (xt ** 2).sum() + 0.05 * xt.abs().sum()
    lt.backward()
    assert abs(loss - float(lt)) <= 1e-5 * abs(float(lt))
    np.testing.assert_allclose(g, xt.grad.numpy(), atol=1e-5)
