"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Gates from BASELINE.json north_star: HR maps within 1e-4 max-abs (fp32), masks >= 99.9 % identical,
IoU within 0.1 points.  Because two honest fp32 evaluations of this solver already drift by ~7e-4
through sign()/Adam (profiles/r01_chaos_floor.txt), the kernels reproduce the oracle's un-fused
evaluation order and the tests below assert the stronger property: BIT-IDENTICAL maps.  The loss is a
reported scalar summed in a different order, so it is compared to 1e-5 relative.
Nothing here reads /root/reference.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import golden_params, load_golden

pytestmark = pytest.mark.gpu

A = None
O = None


@pytest.fixture(scope="module", autouse=True)
def _mods(oracle):
    global A, O
    from deeplabv3plus_augmented_superresolution_b200 import _lib
    _lib.lib()
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    A, O = _lib, oracle


def synth(B, N, hw, angle_max=0.15, shift_max=80, seed=1234, value=1.0):
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
    h, w = hw
    return make_augmented_copies(B, N, (h, w), (4 * h, 4 * w), angle_max, shift_max, seed, value, device="cuda")


def okw(kw):
    return {k: v for k, v in kw.items() if k != "images_in_flight"}


def assert_loss_close(a, b):
    assert abs(a - b) <= 1e-5 * max(1.0, abs(b)), (a, b)


# --------------------------------------------------------------------------------------------------
# the solve
# --------------------------------------------------------------------------------------------------
CASES = [
    dict(N=6, hw=(32, 32), iters=20),
    dict(N=5, hw=(32, 32), iters=10, angle_max=3.1, shift_max=30, seed=7),                      # any rotation
    dict(N=4, hw=(16, 16), iters=10, angle_max=0.5, shift_max=70, seed=3, value=8.0),           # test_SR.py's {0,8} maps, shifts > canvas
    dict(N=3, hw=(24, 40), iters=8, angle_max=0.3, shift_max=25, seed=9),                       # ragged: partial tiles, h != w
    dict(N=1, hw=(16, 16), iters=5, seed=4),                                                    # single (identity) copy
    dict(N=7, hw=(32, 32), iters=12, seed=5, kw=dict(lambda_l1=0.05, amsgrad=False, step_offset=600)),
    dict(N=5, hw=(16, 16), iters=15, seed=2, kw=dict(optimizer="sgd", learning_rate=1e-4, lr_scheduler=False)),
    dict(N=5, hw=(16, 16), iters=15, seed=2, kw=dict(optimizer="sgd", momentum=0.9, nesterov=True, learning_rate=1e-4)),
    dict(N=5, hw=(16, 16), iters=15, seed=2, kw=dict(optimizer="sgd", momentum=0.5, learning_rate=1e-4, lr_scheduler=False)),
    dict(N=5, hw=(16, 16), iters=15, seed=6, kw=dict(optimizer="adagrad", learning_rate=1e-2, epsilon=1e-7)),
    dict(N=5, hw=(16, 16), iters=15, seed=6, kw=dict(optimizer="adadelta", learning_rate=1.0)),
    dict(N=5, hw=(16, 16), iters=15, seed=6, kw=dict(optimizer="adamax", learning_rate=2e-3, step_offset=45)),
    dict(N=4, hw=(16, 16), iters=6, seed=8, kw=dict(lambda_df=0.37, lambda_tv=0.11, lambda_l2=0.0, decay_steps=3, decay_rate=0.5)),
    dict(N=5, hw=(32, 32), iters=12, seed=12, kw=dict(use_btv=True)),                                       # bilateral TV
    dict(N=4, hw=(16, 24), iters=9, seed=13, value=8.0, kw=dict(use_btv=True, lambda_tv=0.05, lambda_l1=0.01, amsgrad=False)),
    dict(N=5, hw=(18, 22), iters=7, angle_max=0.4, shift_max=15, seed=14),                      # LR width not a multiple of 4: padded residual pitch
    dict(N=4, hw=(15, 17), iters=6, angle_max=1.2, shift_max=10, seed=15, value=8.0),           # odd sizes
    dict(N=131, hw=(16, 16), iters=3, angle_max=0.8, shift_max=30, seed=16),                    # odd copy count over two K2 chunks
    dict(N=5, hw=(256, 256), iters=3, angle_max=0.15, shift_max=160, seed=17),                  # 1024^2 canvas: 256 forward tiles, 256 gradient tiles
    dict(N=4, hw=(200, 72), iters=3, angle_max=1.0, shift_max=40, seed=18),                     # tall ragged canvas, big-box variant
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items() if k != "kw") + "".join(f"-{k}" for k in c.get("kw", {})))
def test_solve_bit_identical_to_oracle(case):
    kw = case.get("kw", {})
    copies, ang, sh = synth(2, case["N"], case["hw"], case.get("angle_max", 0.15), case.get("shift_max", 80), case.get("seed", 1), case.get("value", 1.0))
    h, w = case["hw"]
    x, loss = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=case["iters"], **kw), want_loss=True)
    cp = copies.cpu().numpy()
    for b in range(2):
        xo, lo = O.augmented_superresolution(cp[b], ang[b], sh[b], O.SolveParams(num_iter=case["iters"], **kw), output_size=(4 * h, 4 * w))
        np.testing.assert_array_equal(x[b].cpu().numpy(), xo[..., 0])
        assert_loss_close(float(loss[b]), lo)


@pytest.mark.parametrize("amax", [0.15, 0.1501, 0.152, 0.19, 0.2])
def test_forward_box_variant_boundary(amax):
    """K1 stages a 74-row source box when every |angle| <= ~0.15 rad (the reference's ANGLE_MAX, test_SR.py:31) and a 92-row box otherwise;
    a box that does not fit traps in k_forward_tables.  Copies at exactly +-amax, with shifts that push tiles over every canvas edge."""
    N, h, w, iters = 6, 48, 32, 5
    copies, ang, sh = synth(1, N, (h, w), amax, 60, seed=61)
    ang = np.array([[amax, -amax, np.float32(amax), -np.float32(amax), 0.5 * amax, 0.0]], np.float32)
    x = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=iters))
    xo, _ = O.augmented_superresolution(copies[0].cpu().numpy(), ang[0], sh[0], O.SolveParams(num_iter=iters), output_size=(4 * h, 4 * w))
    np.testing.assert_array_equal(x[0].cpu().numpy(), xo[..., 0])


@pytest.mark.parametrize("B", [1, 6])   # 32-row gradient tiles (K1 lets K2 in early) and 64-row tiles
def test_launch_chaining_modes_are_identical(B, monkeypatch):
    """The two solve kernels are chained with programmatic dependent launch (ASR_PDL bit 0: forward-residual launches, bit 1: gradient
    launches).  Every mode must give the bits of the plain launches, with and without a loss trace between the kernels; a missing wait
    would show up as a race between one kernel's tail and the next one's first wave."""
    copies, ang, sh = synth(B, 24, (64, 64), 0.15, 40, seed=71)
    outs = {}
    for mask in ("0", "1", "2", "3"):
        monkeypatch.setenv("ASR_PDL", mask)
        for rep in range(2):
            x = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=40))
            xt, _, trace = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=40), want_loss=True, loss_every=10)
            outs[(mask, rep)] = (x.cpu().numpy(), xt.cpu().numpy(), trace.cpu().numpy())
    ref = outs[("0", 0)]
    for key, got in outs.items():
        for a, b in zip(ref, got):
            np.testing.assert_array_equal(a, b, err_msg=f"ASR_PDL={key[0]} repetition {key[1]}")
    xo, _ = O.augmented_superresolution(copies[0].cpu().numpy(), ang[0], sh[0], O.SolveParams(num_iter=40), output_size=(256, 256))
    np.testing.assert_array_equal(ref[0][0], xo[..., 0])


@pytest.mark.parametrize("ty", ["32", "64"])
def test_both_gradient_tile_heights(ty, monkeypatch):
    """K2 runs 64x32 tiles for one or two images and 64x64 tiles otherwise; ASR_K2_TY forces either on the same input"""
    monkeypatch.setenv("ASR_K2_TY", ty)
    copies, ang, sh = synth(3, 9, (32, 48), 0.7, 30, seed=55)
    x = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=7))
    for b in range(3):
        xo, _ = O.augmented_superresolution(copies[b].cpu().numpy(), ang[b], sh[b], O.SolveParams(num_iter=7), output_size=(128, 192))
        np.testing.assert_array_equal(x[b].cpu().numpy(), xo[..., 0])


def test_single_evaluation_residual_and_gradient():
    copies, ang, sh = synth(2, 9, (32, 32), 0.6, 40, seed=21)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.rand((2, 128, 128), device="cuda", generator=g) * 1.4 - 0.2).contiguous()
    for kw in (dict(), dict(lambda_l1=0.3, lambda_df=0.8)):
        r, gr, l = A.loss_grad_batched(x, copies, ang, sh, A.SolveParams(**kw))
        for b in range(2):
            lo, go, ro = O.loss_and_grad(x[b].cpu().numpy(), copies[b].cpu().numpy(), ang[b], sh[b], O.SolveParams(**kw), want_resid=True)
            np.testing.assert_array_equal(r[b].cpu().numpy(), ro)
            np.testing.assert_array_equal(gr[b].cpu().numpy(), go)
            assert_loss_close(float(l[b]), lo)
    # with x == 0 the rotated image is exactly zero: r == -y
    r, _, _ = A.loss_grad_batched(torch.zeros_like(x), copies, ang, sh, A.SolveParams())
    np.testing.assert_array_equal(r.cpu().numpy(), -copies.cpu().numpy())


def test_copy_dropout_keep_mask():
    copies, ang, sh = synth(1, 8, (16, 16), 0.3, 20, seed=31)
    keep = np.array([[1, 0, 1, 1, 0, 1, 1, 0]], np.uint8)
    x = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=12), keep=keep)
    xo, _ = O.augmented_superresolution(copies[0].cpu().numpy(), ang[0], sh[0], O.SolveParams(num_iter=12), output_size=(64, 64), keep=keep[0])
    np.testing.assert_array_equal(x[0].cpu().numpy(), xo[..., 0])
    r, g, _ = A.loss_grad_batched(x, copies, ang, sh, A.SolveParams(), keep=keep)
    assert (r[0, [1, 4, 7]] == 0).all()


def test_per_image_parameters_in_one_batch():
    """config 5 (sweep): every image carries its own lambdas / lr / iterations / step offset."""
    copies, ang, sh = synth(3, 5, (16, 16), 0.2, 15, seed=41)
    plist = [dict(num_iter=9, lambda_tv=0.1), dict(num_iter=4, learning_rate=3e-3, step_offset=300), dict(num_iter=13, lambda_l2=0.2, amsgrad=False)]
    x, loss = A.solve_batched(copies, ang, sh, [A.SolveParams(**p) for p in plist], want_loss=True)
    for b, p in enumerate(plist):
        xo, lo = O.augmented_superresolution(copies[b].cpu().numpy(), ang[b], sh[b], O.SolveParams(**p), output_size=(64, 64))
        np.testing.assert_array_equal(x[b].cpu().numpy(), xo[..., 0])
        assert_loss_close(float(loss[b]), lo)


def test_sweep_points_share_stacks():
    """config 5: a hyper-parameter grid over a few images; every (image, point) pair is an independent solve
    reading the image's LR stack in place."""
    copies, ang, sh = synth(2, 6, (16, 16), 0.2, 12, seed=71)
    grid = [dict(lambda_tv=tv, learning_rate=lr, num_iter=it) for tv in (0.1, 0.3) for lr in (1e-3, 3e-3) for it in (5, 8)]
    points = [(s, g) for s in range(2) for g in grid]
    x, loss = A.solve_sweep(copies, ang, sh, [A.SolveParams(**g) for _, g in points], [s for s, _ in points], want_loss=True)
    for i, (s, g) in enumerate(points):
        xo, lo = O.augmented_superresolution(copies[s].cpu().numpy(), ang[s], sh[s], O.SolveParams(**g), output_size=(64, 64))
        np.testing.assert_array_equal(x[i].cpu().numpy(), xo[..., 0])
        assert_loss_close(float(loss[i]), lo)


def test_images_in_flight_groups():
    """launch groups (images_in_flight) address residuals and tap tables by absolute image index"""
    copies, ang, sh = synth(5, 6, (16, 20), 0.3, 20, seed=91)
    x = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=6, images_in_flight=2))
    x1 = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=6))
    np.testing.assert_array_equal(x.cpu().numpy(), x1.cpu().numpy())
    for b in (0, 3, 4):
        xo, _ = O.augmented_superresolution(copies[b].cpu().numpy(), ang[b], sh[b], O.SolveParams(num_iter=6), output_size=(64, 80))
        np.testing.assert_array_equal(x[b].cpu().numpy(), xo[..., 0])


@pytest.mark.parametrize("n_aug", [16, 1024])
def test_num_aug_extremes(n_aug):
    """config 4: the copy count is a free axis (16 ... 1024); more than one 128-copy chunk in K2."""
    copies, ang, sh = synth(1, n_aug, (16, 16), 0.15, 20, seed=81)
    x = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=4))
    xo, _ = O.augmented_superresolution(copies[0].cpu().numpy(), ang[0], sh[0], O.SolveParams(num_iter=4), output_size=(64, 64))
    np.testing.assert_array_equal(x[0].cpu().numpy(), xo[..., 0])


@pytest.mark.parametrize("name", ["small_adam", "small_value8_bigangle", "small_sgd", "canonical_config1"])
def test_committed_goldens(name):
    """canonical_config1 = BASELINE.json configs[0]: 1 image, 100 copies, 128^2 -> 512^2, 300 Adam/AMSGrad steps."""
    g = load_golden(name)
    kw = golden_params(g)
    copies = torch.from_numpy(g["copies"]).cuda()[None].contiguous()
    x, loss = A.solve_batched(copies, g["angles"][None], g["shifts"][None], A.SolveParams(**kw), want_loss=True)
    xg = x[0].cpu().numpy()
    assert np.abs(xg - g["x"]).max() <= 1e-4                       # the stated gate ...
    np.testing.assert_array_equal(xg, g["x"])                      # ... and the one actually met
    assert_loss_close(float(loss[0]), float(g["loss"]))
    for it, ref in zip(g["trace_iters"], g["trace"]):
        kw2 = dict(kw, num_iter=int(it))
        xi = A.solve_batched(copies, g["angles"][None], g["shifts"][None], A.SolveParams(**kw2))
        np.testing.assert_array_equal(xi[0].cpu().numpy(), ref)
    _, gr, l0 = A.loss_grad_batched(torch.from_numpy(O.resize_bilinear(g["copies"][:1, :, :, None], g["x"].shape)[0, :, :, 0]).cuda()[None].contiguous(),
                                    copies, g["angles"][None], g["shifts"][None], A.SolveParams(**kw))
    np.testing.assert_array_equal(gr[0].cpu().numpy(), g["grad0"])
    # thresholded masks and IoU (the other two gates) -- identical maps give identical masks
    for th in (0.2, 0.65):
        mg = O.threshold_image(xg, 8, th_factor=th)
        mo = O.threshold_image(g["x"], 8, th_factor=th)
        assert (mg == mo).mean() >= 0.999
        assert abs(O.compute_iou(mo, mg, 8) - 1.0) <= 1e-3


def test_full_size_batch_properties():
    """BASELINE.json configs[1] shape (128^2 -> 512^2, 100 copies) without a CPU solve: properties that
    hold for any correct implementation of independent per-image solves."""
    B = 6
    copies, ang, sh = synth(B, 100, (128, 128), seed=77)
    copies[3] = copies[0]; ang[3] = ang[0]; sh[3] = sh[0]
    P = A.SolveParams(num_iter=12)
    x = A.solve_batched(copies, ang, sh, P)
    assert torch.isfinite(x).all()
    assert torch.equal(x[0], x[3])                                                   # same input -> same output, any slot
    x_alone = A.solve_batched(copies[2:3].contiguous(), ang[2:3], sh[2:3], P)
    assert torch.equal(x_alone[0], x[2])                                             # batching never couples images
    x_grp = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=12, images_in_flight=4))
    assert torch.equal(x_grp, x)                                                     # launch grouping is invisible
    perm = [5, 1, 4, 0, 2, 3]
    x_perm = A.solve_batched(copies[perm].contiguous(), ang[perm], sh[perm], P)
    assert torch.equal(x_perm, x[perm])


def test_full_occupancy_determinism_and_parity():
    """48 full-size images (3072 gradient CTAs, 422k forward CTAs per launch: every SM holds co-resident CTAs of the
    warp-specialised kernel).  A race in the tile hand-off would show up as run-to-run differences or as a mismatch
    against the oracle; both are checked bit for bit."""
    copies, ang, sh = synth(48, 100, (128, 128), 0.15, 80, seed=97)
    P = A.SolveParams(num_iter=6)
    x1 = A.solve_batched(copies, ang, sh, P)
    x2 = A.solve_batched(copies, ang, sh, P)
    assert torch.equal(x1, x2)
    for b in (0, 23, 47):
        xo, _ = O.augmented_superresolution(copies[b].cpu().numpy(), ang[b], sh[b], O.SolveParams(num_iter=6), output_size=(512, 512))
        np.testing.assert_array_equal(x1[b].cpu().numpy(), xo[..., 0])


# --------------------------------------------------------------------------------------------------
# warp, OPM, normalise, threshold, back-projection
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("interp", ["bilinear", "nearest"])
@pytest.mark.parametrize("shape", [(64, 64, 3), (48, 80, 1), (33, 47, 3)])
def test_warp_affine(interp, shape):
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.augmentation_utils import warp_copies
    rng = np.random.RandomState(3)
    img = rng.rand(*shape).astype(np.float32)
    ang = np.array([0.0, 0.15, -0.4, 2.9, 0.05], np.float32)
    shf = np.array([[0, 0], [7.3, -4.6], [-20.5, 30.0], [0.5, 0.5], [3.0, -2.0]], np.float32)
    out = warp_copies(img, ang, shf, interp).cpu().numpy()
    tiled = np.repeat(img[None], len(ang), 0)
    ref = O.translate(O.rotate(tiled, ang, interp), shf, interp)
    np.testing.assert_array_equal(out, ref)
    np.testing.assert_array_equal(out[0], img)                                      # copy 0 is the identity


def test_create_augmented_copies_uses_global_numpy_rng():
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.augmentation_utils import create_augmented_copies
    img = np.random.RandomState(0).rand(32, 32, 3).astype(np.float32)
    np.random.seed(1234)
    copies, a, s = create_augmented_copies(img, 100, 0.15, 80)
    np.testing.assert_allclose(a[1:3], [0.03663263, -0.018681679], atol=1e-8)
    assert a[0] == 0 and (s[0] == 0).all() and copies.shape == (100, 32, 32, 3)
    np.testing.assert_array_equal(copies[0].cpu().numpy(), img)


@pytest.mark.parametrize("mode", ["argmax", "slice", "slice_max"])
def test_opm_extract(mode):
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.augmentation_utils import extract_opm
    rng = np.random.RandomState(5)
    logits = rng.randn(7, 24, 40, 21).astype(np.float32)
    logits[0, 0, 0, :] = 0.5                       # all-tie pixel -> class 0
    logits[1, 3, 3, 8] = logits[1, 3, 3].max()     # tie between class 8 and an earlier/later channel
    c, m = extract_opm(logits, 8, mode)
    co, mo = O.opm_extract(logits, 8, mode)
    np.testing.assert_array_equal(c.cpu().numpy(), co)
    if mode == "slice_max":
        np.testing.assert_array_equal(m.cpu().numpy(), mo)
    else:
        assert m is None
    if mode == "argmax":
        assert set(np.unique(co)) <= {0.0, 8.0}


def test_minmax_normalize_and_threshold():
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts import superres_utils as SU
    rng = np.random.RandomState(6)
    stack = torch.from_numpy((rng.randn(9, 32, 32, 1) * 3 + 1).astype(np.float32)).cuda()
    np.testing.assert_array_equal(SU._normalize_stack_device(stack).cpu().numpy(), O.minmax_normalize_global(stack.cpu().numpy()))
    const = torch.full((4, 8, 8, 1), 2.5, device="cuda")
    assert (SU._normalize_stack_device(const) == 0).all()                            # max == min -> denominator 1
    x = rng.randn(64, 64, 1).astype(np.float32)
    for th in (0.15, 0.2, 0.65):
        np.testing.assert_array_equal(SU.threshold_image(x, 8, th_factor=th), O.threshold_image(x, 8, th_factor=th))
    m = rng.randn(64, 64, 1).astype(np.float32)
    m[:4] = x[:4]                                                                    # equality must count as >=
    out = SU.threshold_image(x, 15, th_mask=m)
    np.testing.assert_array_equal(out, O.threshold_image(x, 15, th_mask=m))
    assert out.dtype == np.int32 and out.shape == (64, 64, 1) and (out[:4] == 15).all()


@pytest.mark.parametrize("mode", ["max", "mean"])
def test_backproject(mode):
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.optimizer import Optimizer
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution
    copies, ang, sh = synth(2, 9, (32, 32), 0.5, 40, seed=13)
    copies = copies + 0.25 * torch.rand_like(copies)
    s = Superresolution(1, 0.3, 0.7, 0, optimizer=Optimizer(), feature_size=(32, 32), output_size=(128, 128))
    out = s.backproject_batched(copies, ang, sh, mode)
    for b in range(2):
        np.testing.assert_array_equal(out[b].cpu().numpy(), O.backproject(copies[b].cpu().numpy(), ang[b], sh[b], mode, (128, 128))[..., 0])
    fn = s.max_superresolution if mode == "max" else s.mean_superresolution
    one, none = fn([c[..., None] for c in copies[0].cpu().numpy()], ang[0], sh[0])
    assert none is None and one.shape == (128, 128, 1) and one.dtype == np.float32
    np.testing.assert_array_equal(one[..., 0], out[0].cpu().numpy())


def test_iou_counts_on_device():
    from deeplabv3plus_augmented_superresolution_b200 import utils
    rng = np.random.RandomState(8)
    t = rng.choice([0, 8, 15, 255], size=(5, 64, 64, 1), p=[0.6, 0.25, 0.1, 0.05]).astype(np.int32)
    p = np.where(rng.rand(5, 64, 64, 1) < 0.8, np.where(t == 8, 8, 0), rng.choice([0, 8], size=t.shape)).astype(np.int32)
    t[3] = 0; p[3] = 0                      # class absent everywhere: NaN without bg, 1.0 with bg
    for bg in (False, True):
        got = utils.compute_IoU_batched(torch.from_numpy(t).cuda(), torch.from_numpy(p).cuda(), 8, include_bg=bg)
        for b in range(5):
            ref = utils.compute_IoU(t[b], p[b], img_size=(64, 64), class_id=8, include_bg=bg)
            orc = O.compute_iou(t[b], p[b], 8, include_bg=bg)
            assert (np.isnan(ref) and np.isnan(got[b]) and np.isnan(orc)) or (got[b] == ref and abs(orc - ref) < 1e-15)


# --------------------------------------------------------------------------------------------------
# the reference-facing Python surface
# --------------------------------------------------------------------------------------------------
def test_reference_api_end_to_end(tmp_path):
    """test_SR.py / SR_single_class.py call sequence with a stand-in model: warp -> predict -> OPM ->
    hdf5 -> load_SR_data -> compute_SR (aug, max, mean) with one shared Optimizer."""
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.optimizer import Optimizer
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts import superres_utils as SU
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts import augmentation_utils as AU
    from PIL import Image

    rng = np.random.RandomState(0)
    yy, xx = np.mgrid[:128, :128]
    rgb = np.stack([(np.hypot(yy - 60, xx - 70) < 35) * 255, (xx * 2) % 256, (yy * 2) % 256], -1).astype(np.uint8)
    img_path = str(tmp_path / "2007_000032.jpg")
    Image.fromarray(rgb).save(img_path, quality=95)

    class FakeModel:   # upstream producer stand-in: [N,128,128,3] -> [N,32,32,21] logits
        def predict(self, images, batch_size=16):
            x = images if isinstance(images, torch.Tensor) else torch.from_numpy(images).cuda()
            pooled = torch.nn.functional.avg_pool2d(x.permute(0, 3, 1, 2), 4)
            w = torch.from_numpy(rng.randn(21, 3).astype(np.float32)).cuda()
            w[8] = torch.tensor([3.0, -1.0, -1.0])
            return torch.einsum("kc,nchw->nhwk", w, pooled).contiguous()

    np.random.seed(1234)
    cm, mm, ang, sh, name = AU.compute_augmented_feature_maps(img_path, FakeModel(), filter_class_id=8, mode="slice_max", num_aug=12,
                                                              angle_max=0.15, shift_max=20, image_size=(128, 128), dest_folder=str(tmp_path / "h5"))
    assert name == "2007_000032" and len(cm) == 12 and len(mm) == 12 and cm[0].shape == (32, 32, 1)
    class_masks, max_masks, a2, s2, fname = SU.load_SR_data(str(tmp_path / "h5" / "2007_000032.hdf5"), num_aug=10)
    assert fname == name and class_masks.shape == (10, 32, 32, 1) and max_masks.shape == (10, 32, 32, 1)
    np.testing.assert_array_equal(a2, ang[:10]); np.testing.assert_array_equal(s2, sh[:10])
    np.testing.assert_array_equal(class_masks.cpu().numpy(), O.minmax_normalize_global(np.stack(cm[:10])))
    with pytest.raises(Exception, match="is invalid"):
        SU.load_SR_data(str(tmp_path / "h5" / "2007_000032.hdf5"), num_aug=13)

    opt = Optimizer(optimizer="adam", learning_rate=1e-3, amsgrad=True, lr_scheduler=True, decay_steps=60, decay_rate=0.3)
    sr = Superresolution(lambda_df=1.0, lambda_tv=0.3, lambda_L2=0.7, lambda_L1=0.0, num_iter=15, num_aug=10, optimizer=opt,
                         feature_size=(32, 32), output_size=(128, 128))
    cmn, mmn = class_masks.cpu().numpy(), max_masks.cpu().numpy()
    P = lambda off: O.SolveParams(num_iter=15, step_offset=off)
    # aug: class solve then max solve on the SAME optimizer -> step offsets 0 and 15 (Appendix B-1)
    th = SU.compute_SR(sr, class_masks, a2, s2, fname, str(tmp_path / "out"), SR_type="aug", max_masks=max_masks, class_id=8)
    xc, _ = O.augmented_superresolution(cmn, a2, s2, P(0), output_size=(128, 128))
    xm, _ = O.augmented_superresolution(mmn, a2, s2, P(15), output_size=(128, 128))
    np.testing.assert_array_equal(th, O.threshold_image(xc, 8, th_mask=xm))
    assert th.dtype == np.int32 and th.shape == (128, 128, 1) and opt.iterations == 30
    assert (tmp_path / "out" / "aug_SR").is_dir()
    # argmax-style call: max_masks None must behave like [] (the reference crashes, Appendix B-2)
    th2 = SU.compute_SR(sr, class_masks, a2, s2, fname, str(tmp_path / "out"), SR_type="aug", max_masks=None, class_id=8, th_factor=0.65,
                        save_final_output=True)
    xc2, _ = O.augmented_superresolution(cmn, a2, s2, P(30), output_size=(128, 128))
    np.testing.assert_array_equal(th2, O.threshold_image(xc2, 8, th_factor=0.65))
    assert (tmp_path / "out" / "aug_SR" / f"{fname}_aug_SR.png").exists()
    for kind in ("max", "mean"):
        t = SU.compute_SR(sr, class_masks, a2, s2, fname, str(tmp_path / "out"), SR_type=kind, max_masks=[], class_id=8, th_factor=0.2)
        np.testing.assert_array_equal(t, O.threshold_image(O.backproject(cmn, a2, s2, kind, (128, 128)), 8, th_factor=0.2))
    # list-of-arrays input (test_SR.py path) and the returned tuple
    x, loss = sr.augmented_superresolution([c for c in cmn], a2, s2)
    assert isinstance(x, np.ndarray) and x.shape == (128, 128, 1) and x.dtype == np.float32 and isinstance(loss, float)
    xo, lo = O.augmented_superresolution(cmn, a2, s2, P(45), output_size=(128, 128))
    np.testing.assert_array_equal(x, xo)
    assert_loss_close(loss, lo)


def test_batched_api_mirrors_sequential_shared_optimizer():
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.optimizer import Optimizer
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution
    copies, ang, sh = synth(3, 6, (16, 16), 0.2, 10, seed=55)
    mk = lambda: Superresolution(1.0, 0.3, 0.7, 0.0, num_iter=7, num_aug=6, optimizer=Optimizer(amsgrad=True, lr_scheduler=True, decay_steps=60, decay_rate=0.3),
                                 feature_size=(16, 16), output_size=(64, 64))
    s1, s2 = mk(), mk()
    xb = s1.augmented_superresolution_batched(copies, ang, sh)
    for b in range(3):
        xs, _ = s2.augmented_superresolution(copies[b].cpu().numpy()[..., None], ang[b], sh[b])
        np.testing.assert_array_equal(xb[b].cpu().numpy(), xs[..., 0])
    assert s1.optimizer.iterations == s2.optimizer.iterations == 21


def test_dlpack_entry_point():
    from deeplabv3plus_augmented_superresolution_b200 import _lib
    copies, ang, sh = synth(2, 5, (16, 16), 0.2, 10, seed=66)
    P = _lib.SolveParams(num_iter=6)
    ref = _lib.solve_batched(copies, ang, sh, P)
    L = _lib.lib()
    need = C.c_size_t()
    _lib.check(L.asr_solve_workspace_bytes(2, 5, 16, 16, 64, 64, 6, C.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
    x = torch.empty((2, 64, 64), dtype=torch.float32, device="cuda")
    caps = [t.__dlpack__() for t in (copies, x, ws)]
    C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
    C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    ptrs = [C.pythonapi.PyCapsule_GetPointer(c, b"dltensor") for c in caps]
    arr, n = _lib._params_array(P)
    a = np.ascontiguousarray(ang, np.float32); s = np.ascontiguousarray(sh, np.float32)
    _lib.check(L.asr_solve_batched_dlpack(arr, n, ptrs[0], a.ctypes.data_as(_lib._fp), s.ctypes.data_as(_lib._fp), None,
                                          ptrs[1], None, ptrs[2], C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert torch.equal(x, ref)
    # a CPU tensor is rejected with ASR_EDTYPE, not silently computed elsewhere
    cpu_cap = copies.cpu().__dlpack__()
    rc = L.asr_solve_batched_dlpack(arr, n, C.pythonapi.PyCapsule_GetPointer(cpu_cap, b"dltensor"), a.ctypes.data_as(_lib._fp),
                                    s.ctypes.data_as(_lib._fp), None, ptrs[1], None, ptrs[2], None)
    assert rc == -6


# --------------------------------------------------------------------------------------------------
# one process, several GPUs: kernel attributes (dynamic shared memory limits) and the SM count are per device
# --------------------------------------------------------------------------------------------------
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_same_process_second_device():
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution
    P = A.SolveParams(num_iter=6)
    sr = Superresolution(1.0, 0.3, 0.7, 0.0, feature_size=(32, 32), output_size=(128, 128))
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        copies, ang, sh = make_augmented_copies(1, 5, (32, 32), (128, 128), 0.15, 20, seed=21, device=dev)
        x = A.solve_batched(copies, ang, sh, P)
        mx = sr.backproject_batched(copies, ang, sh, "max")
        torch.cuda.synchronize(dev)
        outs.append((x.cpu().numpy(), mx.cpu().numpy()))
    xo, _ = O.augmented_superresolution(copies[0].cpu().numpy(), ang[0], sh[0], O.SolveParams(num_iter=6), output_size=(128, 128))
    for x, mx in outs:
        assert np.array_equal(x[0], xo[..., 0])
        assert np.array_equal(mx, outs[0][1])


def test_misaligned_pointers_are_rejected():
    copies, ang, sh = synth(1, 3, (16, 16))
    L = A.lib()
    need = C.c_size_t()
    A.check(L.asr_solve_workspace_bytes(1, 3, 16, 16, 64, 64, 2, C.byref(need)))
    ws = torch.empty(need.value + 512, dtype=torch.uint8, device="cuda")
    x = torch.empty((1, 64, 64), dtype=torch.float32, device="cuda")
    arr, n = A._params_array(A.SolveParams(num_iter=2))
    a32, s32 = np.ascontiguousarray(ang, np.float32), np.ascontiguousarray(sh, np.float32)
    fp = C.POINTER(C.c_float)
    rc = L.asr_solve_batched(arr, n, copies.data_ptr() + 4, a32.ctypes.data_as(fp), s32.ctypes.data_as(fp), None, 1, 3, 16, 16, 64, 64,
                             x.data_ptr(), None, ws.data_ptr(), need.value, None)
    assert rc == -1 and b"aligned" in L.asr_last_error()
    base = (ws.data_ptr() + 255) // 256 * 256
    rc = L.asr_solve_batched(arr, n, copies.data_ptr(), a32.ctypes.data_as(fp), s32.ctypes.data_as(fp), None, 1, 3, 16, 16, 64, 64,
                             x.data_ptr(), None, base + 16, need.value, None)
    assert rc == -1 and b"aligned" in L.asr_last_error()


# --------------------------------------------------------------------------------------------------
# paths VERDICT r01 listed as untested: generic back-projection, per-image normalisation, chunked copies, workspace forms
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["max", "mean"])
@pytest.mark.parametrize("sizes", [((32, 32), (64, 64)), ((24, 40), (72, 120)), ((16, 16), (128, 128))])
def test_backproject_generic_scale(mode, sizes):
    """Superresolution's default feature_size (64,64) -> (512,512) is a x8 back-projection (superresolution.py:28,139-161):
    any output/feature ratio other than 4 takes the one-thread-per-pixel kernel."""
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.optimizer import Optimizer
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution
    (h, w), (H, W) = sizes
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
    copies, ang, sh = make_augmented_copies(2, 7, (h, w), (H, W), 0.4, 0.2 * W, seed=31, device="cuda")
    copies = copies + 0.25 * torch.rand_like(copies)
    s = Superresolution(1, 0.3, 0.7, 0, optimizer=Optimizer(), feature_size=(h, w), output_size=(H, W))
    out = s.backproject_batched(copies, ang, sh, mode)
    for b in range(2):
        np.testing.assert_array_equal(out[b].cpu().numpy(), O.backproject(copies[b].cpu().numpy(), ang[b], sh[b], mode, (H, W))[..., 0])


def test_load_sr_data_per_image_normalisation(tmp_path):
    """global_normalize=False: every copy is normalised with its own min/max (superres_utils.py:56-62 with global_min=None)."""
    from deeplabv3plus_augmented_superresolution_b200 import hdf5_lite
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts import superres_utils as SU
    rng = np.random.RandomState(9)
    cm = [(rng.randn(16, 16, 1) * (k + 1) + k).astype(np.float32) for k in range(6)]
    cm[3][:] = 2.0                                                    # constant copy: max == min -> denominator 1 -> all zeros
    mm = [(rng.rand(16, 16, 1) * 5).astype(np.float32) for _ in range(6)]
    f = hdf5_lite.File(str(tmp_path / "2007_000033.hdf5"), "w")
    f.create_dataset("class_masks", data=cm); f.create_dataset("max_masks", data=mm)
    f.create_dataset("angles", data=np.zeros(6, np.float32)); f.create_dataset("shifts", data=np.zeros((6, 2), np.float32))
    f.attrs["filename"] = "2007_000033"; f.attrs["mode"] = "slice_max"; f.attrs["angle_max"] = 0.15; f.attrs["shift_max"] = 80
    f.close()
    c, m, a, s, name = SU.load_SR_data(str(tmp_path / "2007_000033.hdf5"), num_aug=5, global_normalize=False)
    assert c.shape == (5, 16, 16, 1) and m.shape == (5, 16, 16, 1) and name == "2007_000033"
    for k in range(5):
        np.testing.assert_array_equal(c[k].cpu().numpy(), SU.min_max_normalization(cm[k], 0.0, 1.0).astype(np.float32))
        np.testing.assert_array_equal(c[k].cpu().numpy(), O.minmax_normalize_global(cm[k][None])[0])
        np.testing.assert_array_equal(m[k].cpu().numpy(), O.minmax_normalize_global(mm[k][None])[0])
    assert (c[3] == 0).all()
    cg, _, _, _, _ = SU.load_SR_data(str(tmp_path / "2007_000033.hdf5"), num_aug=5, global_normalize=True)
    np.testing.assert_array_equal(cg.cpu().numpy(), O.minmax_normalize_global(np.stack(cm[:5])))
    # slice mode: no normalisation at all (superres_utils.py:186)
    f = hdf5_lite.File(str(tmp_path / "2007_000034.hdf5"), "w")
    f.create_dataset("class_masks", data=cm); f.create_dataset("angles", data=np.zeros(6, np.float32))
    f.create_dataset("shifts", data=np.zeros((6, 2), np.float32))
    f.attrs["filename"] = "2007_000034"; f.attrs["mode"] = "slice"; f.attrs["angle_max"] = 0.15; f.attrs["shift_max"] = 80
    f.close()
    c2, m2, _, _, _ = SU.load_SR_data(str(tmp_path / "2007_000034.hdf5"), num_aug=6)
    assert m2 is None
    np.testing.assert_array_equal(c2.cpu().numpy(), np.stack(cm))


def test_create_augmented_copies_chunked():
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts import augmentation_utils as AU
    img = np.random.RandomState(1).rand(40, 48, 3).astype(np.float32)
    np.random.seed(77)
    whole, a1, s1 = AU.create_augmented_copies(img, 12, 0.3, 10)
    np.random.seed(77)
    chunks, a2, s2 = AU.create_augmented_copies_chunked(img, 12, 0.3, 10, chunk_size=4)
    assert isinstance(chunks, np.ndarray) and chunks.shape == (12, 40, 48, 3)
    np.testing.assert_array_equal(a1, a2); np.testing.assert_array_equal(s1, s2)
    np.testing.assert_array_equal(whole.cpu().numpy(), chunks)
    with pytest.raises(Exception, match="multiple"):
        AU.create_augmented_copies_chunked(img, 10, 0.3, 10, chunk_size=4)


def test_workspace_forms_of_aux_calls():
    """asr_warp_affine_ws / asr_backproject_batched_ws give the results of the allocating forms and check their workspace."""
    L = A.lib()
    fp = C.POINTER(C.c_float)
    img = torch.rand((48, 64, 3), device="cuda")
    ang = np.array([0.0, 0.2, -0.1], np.float32); shf = np.array([[0, 0], [5.5, -3], [-8, 2.25]], np.float32)
    need = C.c_size_t()
    A.check(L.asr_warp_affine_workspace_bytes(3, 48, 64, 3, C.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
    o1 = torch.empty((3, 48, 64, 3), device="cuda"); o2 = torch.empty_like(o1)
    A.check(L.asr_warp_affine_ws(img.data_ptr(), ang.ctypes.data_as(fp), shf.ctypes.data_as(fp), 3, 48, 64, 3, 1, o1.data_ptr(), ws.data_ptr(), need.value, None))
    A.check(L.asr_warp_affine(img.data_ptr(), ang.ctypes.data_as(fp), shf.ctypes.data_as(fp), 3, 48, 64, 3, 1, o2.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)
    assert L.asr_warp_affine_ws(img.data_ptr(), ang.ctypes.data_as(fp), shf.ctypes.data_as(fp), 3, 48, 64, 3, 1, o1.data_ptr(), ws.data_ptr(), need.value - 1, None) == -5
    copies, a2, s2 = synth(2, 5, (16, 16), 0.3, 8, seed=41)
    a32, s32 = np.ascontiguousarray(a2, np.float32), np.ascontiguousarray(s2, np.float32)
    A.check(L.asr_backproject_workspace_bytes(2, 5, C.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
    b1 = torch.empty((2, 64, 64), device="cuda"); b2 = torch.empty_like(b1)
    A.check(L.asr_backproject_batched_ws(1, copies.data_ptr(), a32.ctypes.data_as(fp), s32.ctypes.data_as(fp), 2, 5, 16, 16, 64, 64, b1.data_ptr(), ws.data_ptr(), need.value, None))
    A.check(L.asr_backproject_batched(1, copies.data_ptr(), a32.ctypes.data_as(fp), s32.ctypes.data_as(fp), 2, 5, 16, 16, 64, 64, b2.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.equal(b1, b2)
    assert L.asr_backproject_batched_ws(1, copies.data_ptr(), a32.ctypes.data_as(fp), s32.ctypes.data_as(fp), 2, 5, 16, 16, 64, 64, b1.data_ptr(), ws.data_ptr(), 8, None) == -5


# --------------------------------------------------------------------------------------------------
# output/feature ratios other than 4 (Superresolution's default feature_size (64,64) -> (512,512) is x8)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [
    dict(hw=(16, 16), S=8, N=5, iters=6, angle_max=0.15, shift_max=20, seed=51),
    dict(hw=(24, 20), S=2, N=4, iters=8, angle_max=0.6, shift_max=6, seed=52, value=8.0),
    dict(hw=(12, 16), S=6, N=3, iters=5, angle_max=2.0, shift_max=15, seed=53, kw=dict(optimizer="sgd", momentum=0.9, learning_rate=1e-4)),
    dict(hw=(16, 16), S=8, N=4, iters=5, angle_max=0.3, shift_max=10, seed=54, kw=dict(use_btv=True, lambda_l1=0.02)),
], ids=lambda c: f"x{c['S']}-{c['hw'][0]}x{c['hw'][1]}")
def test_solve_other_even_ratios(case):
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
    h, w = case["hw"]; S = case["S"]; H, W = S * h, S * w
    kw = case.get("kw", {})
    copies, ang, sh = make_augmented_copies(2, case["N"], (h, w), (H, W), case["angle_max"], case["shift_max"], seed=case["seed"],
                                            value=case.get("value", 1.0), device="cuda")
    x, loss = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=case["iters"], **kw), want_loss=True, output_size=(H, W))
    cp = copies.cpu().numpy()
    for b in range(2):
        xo, lo = O.augmented_superresolution(cp[b], ang[b], sh[b], O.SolveParams(num_iter=case["iters"], **kw), output_size=(H, W))
        assert np.array_equal(x[b].cpu().numpy(), xo[..., 0]), np.abs(x[b].cpu().numpy() - xo[..., 0]).max()
        assert_loss_close(float(loss[b]), lo)
    # single evaluation: residual and gradient
    xs = torch.rand((2, H, W), device="cuda")
    resid, grad, _ = A.loss_grad_batched(xs, copies, ang, sh, A.SolveParams(**kw))
    for b in range(2):
        _, g, r = O.loss_and_grad(xs[b].cpu().numpy(), cp[b], ang[b], sh[b], O.SolveParams(**kw), want_resid=True)
        np.testing.assert_array_equal(resid[b].cpu().numpy(), r)
        np.testing.assert_array_equal(grad[b].cpu().numpy(), g)


def test_reference_class_default_feature_size_is_x8():
    """Superresolution() without feature_size (64,64 -> 512,512: superresolution.py:28) goes through the reference-named call."""
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.optimizer import Optimizer
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
    copies, ang, sh = make_augmented_copies(1, 4, (64, 64), (512, 512), 0.15, 40, seed=61)
    sr = Superresolution(1.0, 0.3, 0.7, 0.0, num_iter=3, num_aug=4, optimizer=Optimizer(amsgrad=True, lr_scheduler=True, decay_steps=60, decay_rate=0.3))
    assert sr.feature_size == (64, 64) and sr.output_size == (512, 512)
    x, loss = sr.augmented_superresolution([c[..., None] for c in copies[0].numpy()], ang[0], sh[0])
    xo, lo = O.augmented_superresolution(copies[0].numpy(), ang[0], sh[0], O.SolveParams(num_iter=3), output_size=(512, 512))
    np.testing.assert_array_equal(x, xo)
    with pytest.raises(NotImplementedError):
        Superresolution(1.0, 0.3, 0.7, 0.0, optimizer=Optimizer(), output_size=(96, 96))._check_sizes(32, 32)       # x3: odd ratio


# --------------------------------------------------------------------------------------------------
# verbose trace: the loss the reference prints every tenth iteration (superresolution.py:129-130)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S", [4, 8])
def test_loss_trace_matches_oracle(S, capsys):
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.optimizer import Optimizer
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution
    h = w = 16
    copies, ang, sh = make_augmented_copies(2, 5, (h, w), (S * h, S * w), 0.2, 10, seed=71, device="cuda")
    plist = [A.SolveParams(num_iter=25), A.SolveParams(num_iter=12, lambda_tv=0.1)]           # per-image iteration counts
    x, loss, trace = A.solve_batched(copies, ang, sh, plist, want_loss=True, output_size=(S * h, S * w), loss_every=10)
    assert trace.shape == (2, 3)
    tr = trace.cpu().numpy()
    cp = copies.cpu().numpy()
    for b, kw in enumerate([dict(), dict(lambda_tv=0.1)]):
        n_it = plist[b].num_iter
        for j in range(3):
            if 10 * j >= n_it:
                assert np.isnan(tr[b, j])                                                   # untouched: the image had finished
                continue
            _, lo = O.augmented_superresolution(cp[b], ang[b], sh[b], O.SolveParams(num_iter=10 * j + 1, **kw), output_size=(S * h, S * w))
            assert_loss_close(float(tr[b, j]), lo)
        xo, lo = O.augmented_superresolution(cp[b], ang[b], sh[b], O.SolveParams(num_iter=n_it, **kw), output_size=(S * h, S * w))
        assert np.array_equal(x[b].cpu().numpy(), xo[..., 0])                               # tracing does not disturb the solve
        assert_loss_close(float(loss[b]), lo)
    # the reference-named call prints every tenth iteration and the last one
    sr = Superresolution(1.0, 0.3, 0.7, 0.0, num_iter=25, num_aug=5, optimizer=Optimizer(amsgrad=True, lr_scheduler=True, decay_steps=60, decay_rate=0.3),
                         feature_size=(h, w), output_size=(S * h, S * w), verbose=True)
    sr.augmented_superresolution(cp[0][..., None], ang[0], sh[0])
    lines = [l for l in capsys.readouterr().out.splitlines() if " -- loss = " in l]
    assert [l.split(" -- ")[0] for l in lines] == ["1/25", "11/25", "21/25", "25/25"]
    assert abs(float(lines[1].split("= ")[1]) - float(tr[0, 1])) <= 1e-5 * abs(float(tr[0, 1]))
