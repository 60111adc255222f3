"""GPU tests of the in-repo counterparts of the reference's entry scripts (SR_single_class.py:49-141, test_SR.py:57-100).

The batched directory run must produce exactly the masks and IoUs of the reference's sequential per-image loop
(load_SR_data -> compute_SR aug/max/mean with ONE shared Optimizer -> compute_IoU), which in turn is pinned bit for bit to the
oracle by tests/test_parity_gpu.py.  Nothing here reads /root/reference.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from deeplabv3plus_augmented_superresolution_b200 import hdf5_lite, utils
from deeplabv3plus_augmented_superresolution_b200.synthetic import SyntheticSegmenter, make_augmented_copies


def _write_dir(tmp_path, n_files, mode, num_aug=8, hw=(32, 32), value=8.0, bad_at=None):
    """A directory as generate_augmented_copies.py writes it (+ ground-truth PNGs): returns (dir, gt_dir, names)."""
    from PIL import Image
    d, gt = tmp_path / f"xception_{mode}_8_{num_aug}_validation", tmp_path / "SegmentationClassAug"
    d.mkdir(); gt.mkdir(exist_ok=True)
    copies, ang, sh = make_augmented_copies(n_files, num_aug, hw, (4 * hw[0], 4 * hw[1]), 0.15, 20, seed=77, value=value)
    rng = np.random.RandomState(3)
    names = []
    for b in range(n_files):
        name = f"2007_{b:06d}"
        names.append(name)
        cm = copies[b].numpy()[..., None]
        if mode == "slice_max":
            cm = cm / value * 3.0 + rng.rand(*cm.shape).astype(np.float32)          # raw class logits
        f = hdf5_lite.File(str(d / f"{name}.hdf5"), "w")
        n_rows = num_aug - 2 if b == bad_at else num_aug
        f.create_dataset("class_masks", data=[c for c in cm[:n_rows]])
        if mode == "slice_max":
            f.create_dataset("max_masks", data=[(1.5 + rng.rand(*c.shape)).astype(np.float32) for c in cm[:n_rows]])
        f.create_dataset("angles", data=ang[b][:n_rows])
        f.create_dataset("shifts", data=sh[b][:n_rows])
        f.attrs["filename"] = name
        f.attrs["mode"] = mode
        f.attrs["angle_max"] = 0.15
        f.attrs["shift_max"] = 20
        f.close()
        # ground truth: the un-augmented copy upsampled by pixel repetition
        g = (np.kron(copies[b, 0].numpy() > 0, np.ones((4, 4))) * 8).astype(np.uint8)
        Image.fromarray(g, mode="L").save(str(gt / f"{name}.png"))
    return str(d), str(gt), names


def _sequential_reference_loop(d, gt, num_aug, hw, num_iter, th_factor, out_dir):
    """SR_single_class.py:83-127 as written: one image at a time through the reference-named functions."""
    from deeplabv3plus_augmented_superresolution_b200 import SR_single_class as E
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts import superres_utils as SU
    H, W = 4 * hw[0], 4 * hw[1]
    sr = E.build_solver(num_aug=num_aug, feature_size=hw, output_size=(H, W), num_iter=num_iter)
    masks, ious = {"aug": [], "max": [], "mean": []}, {"aug_single": [], "aug_multiple": [], "max": [], "mean": []}
    for path in SU.list_precomputed_data_paths(d, sort=True):
        try:
            cm, mm, ang, sh, name = SU.load_SR_data(path, num_aug=num_aug, global_normalize=True)
        except Exception:
            continue
        true = utils.load_image(os.path.join(gt, f"{name}.png"), image_size=(H, W), normalize=False, is_png=True, resize_method="nearest")
        for kind in ("aug", "max", "mean"):
            m = SU.compute_SR(sr, cm, ang, sh, name, out_dir, SR_type=kind, max_masks=mm, class_id=8, th_factor=th_factor)
            masks[kind].append(m[..., 0])
        ious["aug_single"].append(utils.compute_IoU(true, masks["aug"][-1], img_size=(H, W), class_id=8))
        ious["aug_multiple"].append(utils.compute_IoU(true, masks["aug"][-1], img_size=(H, W), class_id=8, include_bg=True))
        ious["max"].append(utils.compute_IoU(true, masks["max"][-1], img_size=(H, W), class_id=8))
        ious["mean"].append(utils.compute_IoU(true, masks["mean"][-1], img_size=(H, W), class_id=8))
    return masks, {k: float(np.mean(v)) for k, v in ious.items()}, sr.optimizer.iterations


@pytest.mark.parametrize("mode", ["argmax", "slice_max"])
def test_batched_directory_run_equals_sequential_loop(tmp_path, mode, capsys):
    from deeplabv3plus_augmented_superresolution_b200 import SR_single_class as E, batch_runner as BR
    hw, num_aug, num_iter, th = (32, 32), 8, 9, 0.65
    d, gt, names = _write_dir(tmp_path, 5, mode, num_aug=num_aug, hw=hw, bad_at=2)
    seq_masks, seq_avg, seq_steps = _sequential_reference_loop(d, gt, num_aug, hw, num_iter, th, str(tmp_path / "out"))
    assert len(seq_masks["aug"]) == 4                                         # file 2 is invalid (too few rows) and skipped

    # batches of 2 -> (2, 2) with the invalid file reported in between
    sr = E.build_solver(num_aug=num_aug, feature_size=hw, output_size=(128, 128), num_iter=num_iter)
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superres_utils import list_precomputed_data_paths
    got = {"aug": [], "max": [], "mean": []}
    skipped, order = [], []
    for res in BR.run_files(sr, list_precomputed_data_paths(d, sort=True), num_aug=num_aug, class_id=8, th_factor=th, batch=2):
        skipped += res.skipped
        order += res.filenames
        for k in got:
            if res.filenames:
                got[k] += list(getattr(res, k).cpu().numpy())
    assert order == [n for i, n in enumerate(names) if i != 2] and len(skipped) == 1 and skipped[0].endswith(f"{names[2]}.hdf5")
    assert sr.optimizer.iterations == seq_steps == (2 if mode == "slice_max" else 1) * 4 * num_iter
    for k in got:
        for a, b in zip(got[k], seq_masks[k]):
            assert a.dtype == np.int32 and np.array_equal(a, b), k

    # the entry point prints the reference's three lines and returns the same averages
    avg = E.run(d, gt, None, num_aug=num_aug, num_samples=None, class_id=8, th_factor=th, batch=3, img_size=(128, 128),
                feature_size=hw, num_iter=num_iter)
    out = capsys.readouterr().out
    assert "is invalid, skipping..." in out and "Avg. Max SR IoUs:" in out and "Avg. Augmented SR IoUs (with bg)" in out
    assert avg["images"] == 4
    for k in ("aug_single", "aug_multiple", "max", "mean"):
        assert avg[k] == pytest.approx(seq_avg[k], abs=1e-12), k
    assert avg["aug_single"] > 0.5


def test_test_sr_entry_point_with_stand_in_model(tmp_path, capsys):
    from deeplabv3plus_augmented_superresolution_b200 import test_SR as T
    img, gtp = T.write_synthetic_inputs(str(tmp_path))
    ious, masks = T.run(img, gtp, SyntheticSegmenter(21, 8), str(tmp_path / "SR_output"), num_aug=16, num_iter=40)
    assert "Aug. SR (argmax OPM) IoU:" in capsys.readouterr().out
    for k in ("aug", "max", "mean"):
        assert masks[k].shape == (512, 512, 1) and masks[k].dtype == np.int32 and set(np.unique(masks[k])) <= {0, 8}
        assert (tmp_path / "SR_output" / f"{k}_SR" / f"test_shape_{k}_SR.png").exists()
    assert ious["aug"] > 0.9 and ious["max"] > 0.8 and ious["mean"] > 0.8
    with pytest.raises(SystemExit):
        T.main(["--image", img, "--gt", gtp])      # no upstream model given


@pytest.mark.parametrize("mode", ["argmax", "slice_max"])
def test_sweep_grid_equals_point_by_point(tmp_path, mode):
    """sweep_script.run_grid (all hyper-parameter points in one pass, LR stacks shared) == the reference's one-point-per-run loop
    (sweep_script.py:88-161) repeated for every point, including each point's own running optimizer step counter."""
    from deeplabv3plus_augmented_superresolution_b200 import sweep_script as S
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts import superres_utils as SU
    hw, num_aug = (32, 32), 6
    d, gt, names = _write_dir(tmp_path, 4, mode, num_aug=num_aug, hw=hw, bad_at=1)
    grid = [dict(num_iter=7, lambda_tv=0.3, lambda_L2=0.7), dict(num_iter=5, learning_rate=5e-4, amsgrad=True, lambda_tv=1.0),
            dict(num_iter=9, optimizer="sgd", momentum=0.6, learning_rate=1e-4, lambda_tv=0.2, lambda_L2=0.3)]
    got = S.run_grid(grid, d, gt, None, num_aug=num_aug, num_samples=None, class_id=8, th_factor=0.65, batch=2, img_size=(128, 128),
                     feature_size=hw, verbose=False)
    assert len(got) == 3
    for cfg, g in zip(grid, got):
        sr = S.build_solver(cfg, num_aug, hw, (128, 128))
        acc = {k: [] for k in ("aug_iou_single", "aug_iou_multiple", "max_iou", "mean_iou")}
        for path in SU.list_precomputed_data_paths(d, sort=True):
            try:
                cm, mm, ang, sh, name = SU.load_SR_data(path, num_aug=num_aug, global_normalize=True)
            except Exception:
                continue
            true = utils.load_image(os.path.join(gt, f"{name}.png"), image_size=(128, 128), normalize=False, is_png=True, resize_method="nearest")
            m = {k: SU.compute_SR(sr, cm, ang, sh, name, str(tmp_path / "out"), SR_type=k, max_masks=mm, class_id=8, th_factor=0.65)
                 for k in ("aug", "max", "mean")}
            acc["aug_iou_single"].append(utils.compute_IoU(true, m["aug"], img_size=(128, 128), class_id=8))
            acc["aug_iou_multiple"].append(utils.compute_IoU(true, m["aug"], img_size=(128, 128), class_id=8, include_bg=True))
            acc["max_iou"].append(utils.compute_IoU(true, m["max"], img_size=(128, 128), class_id=8))
            acc["mean_iou"].append(utils.compute_IoU(true, m["mean"], img_size=(128, 128), class_id=8))
        assert len(acc["max_iou"]) == 3
        for k, v in acc.items():
            assert g[k] == pytest.approx(float(np.mean(v)), abs=1e-12), (cfg, k)
        assert np.isnan(g["standard_iou_single"])          # no standard masks given
    one = S.run_point(grid[1], d, gt, None, num_aug=num_aug, num_samples=None, class_id=8, th_factor=0.65, batch=4, img_size=(128, 128),
                      feature_size=hw, verbose=False)
    assert one == pytest.approx(got[1], nan_ok=True)


def test_generate_then_solve_through_the_entry_points(tmp_path):
    """generate_augmented_copies (stand-in model) -> hdf5 directory -> SR_single_class.run: the whole chain through the in-repo
    counterparts of the reference's scripts, on a tiny synthetic VOC tree."""
    from PIL import Image
    from deeplabv3plus_augmented_superresolution_b200 import generate_augmented_copies as G, SR_single_class as E
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_test_image
    data = tmp_path / "data"
    voc = data / "dataset_root" / "VOCdevkit" / "VOC2012"
    (voc / "JPEGImages").mkdir(parents=True); (voc / "SegmentationClassAug").mkdir(); (data / "augmented_file_lists").mkdir()
    names = ["2007_000027", "2007_000032", "2007_000033"]
    for i, n in enumerate(names):
        img, gt = make_test_image((128, 128), 8, seed=5 + i)
        if i == 1:
            gt = np.zeros_like(gt)                                   # class 8 absent: filtered out by filter_images_by_class
        Image.fromarray((img * 255.0 + 0.5).astype(np.uint8), mode="RGB").save(str(voc / "JPEGImages" / f"{n}.jpg"), quality=100, subsampling=0)
        Image.fromarray(gt[..., 0].astype(np.uint8), mode="L").save(str(voc / "SegmentationClassAug" / f"{n}.png"))
    (data / "augmented_file_lists" / "valaug.txt").write_text("\n".join(reversed(names)) + "\n")
    args = G.build_parser().parse_args(["--num_aug", "8", "--num_samples", "5", "--mode", "argmax", "--angle_max", "0.15", "--shift_max", "10",
                                        "--use_validation", "--class_id", "8", "--data_dir", str(data)])
    written = G.run(args, SyntheticSegmenter(21, 8), image_size=(128, 128), verbose=False)
    assert [os.path.basename(w) for w in written] == ["2007_000027.hdf5", "2007_000033.hdf5"]      # sorted, class-filtered
    f = hdf5_lite.File(written[0], "r")
    assert f["class_masks"].shape == (8, 32, 32, 1) and f.attrs["mode"] == "argmax" and f.attrs["shift_max"] == 10 and f.attrs["filename"] == "2007_000027"
    assert set(np.unique(f["class_masks"][:8])) <= {0.0, 8.0}
    f.close()
    avg = E.run(os.path.dirname(written[0]), str(voc / "SegmentationClassAug"), None, num_aug=8, num_samples=None, class_id=8, th_factor=0.65,
                batch=2, img_size=(128, 128), feature_size=(32, 32), verbose=False, num_iter=40)
    assert avg["images"] == 2 and avg["aug_single"] > 0.85 and avg["max"] > 0.7 and avg["mean"] > 0.7
    with pytest.raises(SystemExit):
        G.main(["--class_id", "8", "--data_dir", str(data)])        # no upstream model given
