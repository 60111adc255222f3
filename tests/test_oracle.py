"""CPU tests of the oracle (oracle/asr_oracle.c): operator semantics against an independent torch
restatement, the reference's few pinned artefacts, and the committed goldens."""
import os

import numpy as np
import pytest
import torch

from conftest import REFERENCE, golden_params, load_golden
from oracle import torch_mirror as TM


def test_rng_stream_matches_reference_demo_values():
    # np.random.seed(1234) (test_SR.py:16-17) then augmentation_utils.py:14-20; values from SURVEY.md section 4
    from deeplabv3plus_augmented_superresolution_b200.synthetic import draw_angles_shifts
    rng = np.random.RandomState(1234)
    a, s = draw_angles_shifts(rng, 100, 0.15, 80)
    assert a.dtype == np.float32 and s.dtype == np.float32 and a[0] == 0 and (s[0] == 0).all()
    np.testing.assert_allclose(a[1:5], [0.03663263, -0.018681679, 0.08560757, 0.08399274], rtol=0, atol=1e-8)
    np.testing.assert_allclose(s[1:4], [[47.49875, 9.241733], [74.533844, -56.454895], [-75.25648, 15.022959]], rtol=0, atol=1e-5)


def test_transform_matrices(oracle):
    t = oracle.rotate_matrix(0.3, 64, 48)
    np.testing.assert_allclose(t, TM.rotate_matrix(0.3, 64, 48).numpy(), atol=1e-5)
    np.testing.assert_array_equal(oracle.translate_matrix(3.5, -2.25), [1, 0, -3.5, 0, 1, 2.25, 0, 0])
    # inverse of a translation is exact; inverse of a rotation is the rotation by -angle about the same centre
    np.testing.assert_array_equal(oracle.invert_transform(oracle.translate_matrix(3.5, -2.25)), [1, 0, 3.5, 0, 1, -2.25, 0, 0])
    np.testing.assert_allclose(oracle.invert_transform(t), oracle.rotate_matrix(-0.3, 64, 48), atol=2e-5)
    np.testing.assert_allclose(oracle.invert_transform(t), TM.invert(TM.rotate_matrix(0.3, 64, 48)).numpy(), atol=2e-5)


@pytest.mark.parametrize("interp", ["bilinear", "nearest"])
def test_projective_transform_vs_torch(oracle, interp):
    rng = np.random.RandomState(0)
    img = rng.rand(1, 40, 56, 1).astype(np.float32)
    for ang, sh in [(0.0, (0.0, 0.0)), (0.15, (7.3, -4.6)), (-1.1, (-20.5, 30.0)), (3.0, (0.5, 0.5))]:
        rot = oracle.rotate(img, [ang], interp)
        out = oracle.translate(rot, [sh], interp)
        ref = TM.transform(TM.transform(torch.from_numpy(img[0, :, :, 0]), TM.rotate_matrix(ang, 40, 56), interp == "nearest"),
                           TM.translate_matrix(*sh), interp == "nearest").numpy()
        if interp == "bilinear":
            np.testing.assert_allclose(out[0, :, :, 0], ref, atol=2e-5)
        else:   # nearest: identical except where fp32/fp64 rounding of the coordinate crosses a .5 boundary
            assert (out[0, :, :, 0] != ref).mean() < 2e-3


def test_identity_transform_is_exact(oracle):
    img = np.random.RandomState(1).rand(2, 16, 16, 3).astype(np.float32)
    np.testing.assert_array_equal(oracle.translate(oracle.rotate(img, [0.0, 0.0]), [[0, 0], [0, 0]]), img)
    # integer translation moves content right/down for positive shifts, zero fill (A.2)
    out = oracle.translate(img, [[3, 2], [3, 2]])
    np.testing.assert_array_equal(out[:, 2:, 3:], img[:, :-2, :-3])
    assert (out[:, :2] == 0).all() and (out[:, :, :3] == 0).all()


def test_resize_is_box_mean_and_matches_torch(oracle):
    rng = np.random.RandomState(2)
    z = rng.rand(1, 64, 64, 1).astype(np.float32)
    d = oracle.resize_bilinear(z, (16, 16))[0, :, :, 0]
    zz = z[0, :, :, 0]
    box = 0.25 * (zz[1::4, 1::4] + zz[1::4, 2::4] + zz[2::4, 1::4] + zz[2::4, 2::4])   # A.3
    np.testing.assert_allclose(d, box, atol=1e-6)
    np.testing.assert_allclose(d, TM.resize(torch.from_numpy(zz), (16, 16)).numpy(), atol=1e-6)
    y = rng.rand(1, 16, 16, 1).astype(np.float32)
    up = oracle.resize_bilinear(y, (64, 64))[0, :, :, 0]
    np.testing.assert_allclose(up, TM.resize(torch.from_numpy(y[0, :, :, 0]), (64, 64)).numpy(), atol=1e-6)
    # gradient of the resize = exact transpose: <D z, g> == <z, D^T g>
    g = rng.rand(1, 16, 16, 1).astype(np.float32)
    dt = oracle.resize_bilinear_grad(g, (64, 64))
    assert abs((d * g[0, :, :, 0]).sum() - (zz * dt[0, :, :, 0]).sum()) < 1e-3
    assert np.count_nonzero(dt) == 4 * 256


@pytest.mark.parametrize("l1", [0.0, 0.05])
def test_loss_and_grad_vs_torch(oracle, l1):
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
    copies, ang, sh = make_augmented_copies(1, 5, (16, 16), (64, 64), 0.4, 12, seed=3)
    rng = np.random.RandomState(4)
    x = (rng.rand(64, 64) * 0.8 + 0.1).astype(np.float32)
    P = oracle.SolveParams(lambda_l1=l1)
    loss, g, r = oracle.loss_and_grad(x, copies[0].numpy(), ang[0], sh[0], P, want_resid=True)
    lt, gt = TM.loss_and_grad(torch.from_numpy(x), copies[0], ang[0], sh[0], 1.0, 0.3, 0.7, l1)
    assert abs(loss - lt) / abs(lt) < 1e-5
    np.testing.assert_allclose(g, gt.numpy(), atol=2e-4, rtol=1e-5)


def test_gradient_is_tensorflow_not_adjoint(oracle):
    """the registered warp gradient differs from the exact transpose (~1% rel-L2)."""
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
    copies, ang, sh = make_augmented_copies(1, 4, (16, 16), (64, 64), 0.5, 10, seed=5)
    x = torch.rand(64, 64, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    xt = x.clone().requires_grad_(True)
    df = 0
    for k in range(4):
        z = TM.transform(TM.transform(xt, TM.rotate_matrix(float(ang[0, k]), 64, 64)), TM.translate_matrix(*map(float, sh[0, k])))
        df = df + ((TM.resize(z, (16, 16)) - copies[0, k].double()) ** 2).sum()
    df.backward()
    P = oracle.SolveParams(lambda_tv=0.0, lambda_l2=0.0)
    _, g = oracle.loss_and_grad(x.float().numpy(), copies[0].numpy(), ang[0], sh[0], P)
    rel = np.linalg.norm(g - xt.grad.numpy()) / np.linalg.norm(xt.grad.numpy())
    assert 1e-3 < rel < 0.1


def test_optimizer_steps_closed_form(oracle):
    # one Adam step from zero slots moves every pixel with g != 0 by lr*sign(g) (bias-corrected), A.7
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
    copies, ang, sh = make_augmented_copies(1, 3, (16, 16), (64, 64), 0.2, 8, seed=6)
    c = copies[0].numpy()
    P = oracle.SolveParams(num_iter=1, lr_scheduler=False, amsgrad=False)
    x1, _ = oracle.augmented_superresolution(c, ang[0], sh[0], P, output_size=(64, 64))
    x0 = oracle.resize_bilinear(c[:1, :, :, None], (64, 64))[0, :, :, 0]
    _, g = oracle.loss_and_grad(x0, c, ang[0], sh[0], P)
    big = np.abs(g) > 0.05          # epsilon=1e-7 vs sqrt(v)=0.0316*|g| is then < 1e-4 relative
    np.testing.assert_allclose((x0 - x1[..., 0])[big], 1e-3 * np.sign(g[big]), rtol=2e-3)
    assert ((x0 - x1[..., 0])[g == 0] == 0).all()
    # shared-optimizer offset (Appendix B-1): at t=301 there is no beta1 warm-up left, step ~ 0.51*lr
    P2 = oracle.SolveParams(num_iter=1, lr_scheduler=False, amsgrad=False, step_offset=300)
    x2, _ = oracle.augmented_superresolution(c, ang[0], sh[0], P2, output_size=(64, 64))
    ratio = np.median(np.abs((x0 - x2[..., 0])[big])) / 1e-3
    assert abs(ratio - np.sqrt(1 - 0.999 ** 301) / (1 - 0.9 ** 301) * 0.1 / np.sqrt(0.001)) < 0.02
    # SGD: x -= lr * g exactly
    P3 = oracle.SolveParams(num_iter=1, optimizer="sgd", lr_scheduler=False, learning_rate=1e-4)
    x3, _ = oracle.augmented_superresolution(c, ang[0], sh[0], P3, output_size=(64, 64))
    np.testing.assert_array_equal(x3[..., 0], x0 - np.float32(1e-4) * g)


def test_threshold_normalize_opm(oracle):
    rng = np.random.RandomState(7)
    x = rng.randn(32, 32, 1).astype(np.float32)
    th = oracle.threshold_image(x, 8, th_factor=0.2)
    np.testing.assert_array_equal(th, np.where(x > x.max() * np.float32(0.2), 8, 0))
    m = rng.randn(32, 32, 1).astype(np.float32)
    np.testing.assert_array_equal(oracle.threshold_image(x, 3, th_mask=m), np.where(x >= m, 3, 0))
    n = oracle.minmax_normalize_global(x)
    assert n.min() == 0.0 and n.max() == 1.0
    const = np.full((4, 4), 2.5, np.float32)
    np.testing.assert_array_equal(oracle.minmax_normalize_global(const), np.zeros((4, 4), np.float32))   # den -> 1.0
    logits = rng.randn(3, 8, 8, 21).astype(np.float32)
    logits[0, 0, 0, :] = 1.0                                                  # tie -> lowest index (class 0)
    c, _ = oracle.opm_extract(logits, 8, "argmax")
    np.testing.assert_array_equal(c[..., 0], np.where(np.argmax(logits, -1) == 8, 8.0, 0.0))
    assert c[0, 0, 0, 0] == 0.0
    c, _ = oracle.opm_extract(logits, 8, "slice")
    for i in range(3):
        lo, hi = logits[i].min(), logits[i].max()
        np.testing.assert_allclose(c[i, ..., 0], (logits[i, ..., 8] - lo) / (hi - lo), atol=1e-6)
    c, mx = oracle.opm_extract(logits, 8, "slice_max")
    np.testing.assert_array_equal(c[..., 0], logits[..., 8])
    np.testing.assert_array_equal(mx[..., 0], np.delete(logits, 8, axis=-1).max(-1))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "test_images")), reason="reference fixtures not mounted")
def test_iou_known_answers_from_reference_fixtures(oracle):
    """The only numeric artefacts the reference ships: three 512^2 masks and the GT of test_SR.py.
    IoUs computed during the survey: aug 0.75695, max 0.65947, mean 0.76557."""
    from deeplabv3plus_augmented_superresolution_b200 import utils
    ti = os.path.join(REFERENCE, "test_images")
    gt = utils.load_image(os.path.join(ti, "test_cat_gt.png"), image_size=(512, 512), normalize=False, is_png=True, resize_method="nearest")
    assert set(np.unique(gt)) == {0.0, 8.0, 255.0}
    for name, iou, count in (("aug", 0.75695, 33904), ("max", 0.65947, 44638), ("mean", 0.76557, 31888)):
        m = utils.load_image(os.path.join(ti, "SR_output", f"{name}_SR", f"test_cat_{name}_SR.png"), normalize=False, is_png=True)
        m = np.where(m > 0, 8, 0)
        assert int((m > 0).sum()) == count
        assert abs(utils.compute_IoU(gt, m, class_id=8) - iou) < 1e-5
        assert abs(oracle.compute_iou(gt, m, 8) - iou) < 1e-5
        assert abs(oracle.compute_iou(gt, m, 8, include_bg=True) - utils.compute_IoU(gt, m, class_id=8, include_bg=True)) < 1e-12


@pytest.mark.parametrize("name", ["small_adam", "small_value8_bigangle", "small_sgd"])
def test_oracle_reproduces_goldens_bitwise(oracle, name):
    """The oracle is deterministic C (no FMA contraction, fixed summation order): the committed goldens
    must be reproduced bit for bit on any x86-64 box, otherwise GPU-vs-golden tests mean nothing."""
    g = load_golden(name)
    kw = golden_params(g)
    P = oracle.SolveParams(**kw)
    h = g["copies"].shape[1]
    res = oracle.augmented_superresolution(g["copies"], g["angles"], g["shifts"], P, output_size=(4 * h, 4 * h),
                                           trace_iters=list(g["trace_iters"]))
    np.testing.assert_array_equal(res[0][..., 0], g["x"])
    assert np.float32(res[1]) == g["loss"]
    if len(g["trace_iters"]):
        np.testing.assert_array_equal(res[2], g["trace"])
    np.testing.assert_array_equal(oracle.backproject(g["copies"], g["angles"], g["shifts"], "max", (4 * h, 4 * h))[..., 0], g["max_sr"])
    np.testing.assert_array_equal(oracle.backproject(g["copies"], g["angles"], g["shifts"], "mean", (4 * h, 4 * h))[..., 0], g["mean_sr"])


def test_chaos_floor_is_documented(oracle):
    """Two honest fp32 evaluations (FMA contraction on/off) of the same solve drift apart through
    sign()/Adam: this is why the CUDA path reproduces the un-fused order bit for bit."""
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
    copies, ang, sh = make_augmented_copies(1, 8, (32, 32), (128, 128), 0.15, 20, seed=11)
    P = oracle.SolveParams(num_iter=30)
    a, _ = oracle.augmented_superresolution(copies[0].numpy(), ang[0], sh[0], P, output_size=(128, 128))
    b, _ = oracle.augmented_superresolution(copies[0].numpy(), ang[0], sh[0], P, output_size=(128, 128), variant="_fma")
    d = np.abs(a - b)
    assert d.max() > 1e-5           # not bit-identical, and amplified far beyond 1 ulp
    assert np.mean(d) < 1e-4        # but statistically the same solution


_TF_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_crosscheck.npz")


@pytest.mark.skipif(not os.path.exists(_TF_GOLDEN), reason="no TensorFlow-generated golden in this repo yet (oracle/tf_crosscheck.py --write-golden; "
                    "TensorFlow cannot be installed here: profiles/r02_tf_install_attempt.log)")
def test_tf_generated_goldens(oracle):
    """The day a maintainer with tensorflow 2.7 + tfa 0.15 runs oracle/tf_crosscheck.py --write-golden, this pins the oracle to it."""
    z = np.load(_TF_GOLDEN)
    c, a, s, H = z["copies"], z["angles"], z["shifts"], int(z["H"])
    if "rot_tf" in z.files:
        for k in range(len(a)):
            np.testing.assert_allclose(oracle.rotate_matrix(a[k], H, H), z["rot_tf"][k], atol=2e-5, rtol=1e-6)
            np.testing.assert_allclose(oracle.invert_transform(z["rot_tf"][k]), z["rot_inv_tf"][k], atol=3e-4, rtol=1e-5)
            np.testing.assert_array_equal(oracle.invert_transform(z["tr_tf"][k]), z["tr_inv_tf"][k])
    x0 = oracle.resize_bilinear(c[:1, :, :, None], (H, H))[0, :, :, 0]
    lo, g = oracle.loss_and_grad(x0, c, a, s, oracle.SolveParams())
    assert abs(lo - float(z["loss_tf"])) <= 1e-5 * abs(lo)
    assert np.linalg.norm(g - z["grad_tf"]) <= 1e-5 * np.linalg.norm(g)
    for n in (1, 10, int(z["iters"])):
        xo, _ = oracle.augmented_superresolution(c, a, s, oracle.SolveParams(num_iter=n), output_size=(H, H))
        xt = z[f"x_tf_{n}"]
        # x itself can only agree to the fp32 chaos floor (profiles/r01_chaos_floor.txt); the north-star gates that survive it:
        assert np.mean(np.abs(xo - xt)) < 1e-4
        m_o, m_t = oracle.threshold_image(xo, 8, th_factor=0.65), oracle.threshold_image(xt, 8, th_factor=0.65)
        assert (m_o == m_t).mean() >= 0.999
