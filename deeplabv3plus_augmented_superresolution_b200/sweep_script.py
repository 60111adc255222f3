"""B200 counterpart of the reference's sweep_script.py (:1-175): Augmented Super-Resolution hyper-parameter tuning.

The reference runs ONE hyper-parameter point per process (a Weights & Biases agent calls main() with a new `wandb.config`
each time, :76-78) and walks the images one at a time.  W&B itself is out of scope; what this module keeps
is the computation: `run_point(config)` is the reference's main() for one configuration and returns the dict it passes to
`wandb.log` (:163-170); `run_grid(configs)` evaluates MANY configurations in one pass over the data -- every (image, point)
pair is an independent solve, all points of an image read the same low-resolution stack in place (asr_solve_sweep, BASELINE
config 5) -- and returns one such dict per point, identical to what `run_point` gives point by point
(tests/test_entry_points_gpu.py::test_sweep_grid_equals_point_by_point).

    python -m deeplabv3plus_augmented_superresolution_b200.sweep_script --data-dir data --grid grid.json
    (grid.json: a list of objects overriding the defaults below, e.g. [{"lambda_tv": 0.3}, {"lambda_tv": 4.75, "learning_rate": 5e-4}])
"""
import argparse
import json
import os

import numpy as np

from . import _lib
from .batch_runner import _threshold_batched, iou_table
from .superresolution_scripts.optimizer import Optimizer
from .superresolution_scripts.superresolution import Superresolution
from .superresolution_scripts.superres_utils import list_precomputed_data_paths, load_SR_data
from .utils import load_image

SEED = 1234

IMG_SIZE = (512, 512)
FEATURE_SIZE = (128, 128)
NUM_AUG = 100
CLASS_ID = 8
NUM_SAMPLES = 500
MODE = "slice"
MODEL_BACKBONE = "xception"
USE_VALIDATION = False
TH_FACTOR = 0.65

HYPERPARAMETERS_DEFAULT = {          # sweep_script.py:51-73
    "lambda_df": 1,
    "lambda_tv": 4.75,
    "lambda_L2": 0.11,
    "lambda_L1": 0.0,
    "num_iter": 300,
    "use_BTV": False,
    "copy_dropout": 0.0,
    "optimizer": "adam",
    "learning_rate": 1e-3,
    "beta_1": 0.9,
    "beta_2": 0.999,
    "epsilon": 1e-7,
    "amsgrad": False,
    "initial_accumulator_value": 0.1,
    "momentum": 0.6,
    "nesterov": False,
    "lr_scheduler": True,
    "decay_steps": 50,
    "decay_rate": 0.5,
}
LOG_KEYS = ("aug_iou_single", "aug_iou_multiple", "standard_iou_single", "standard_iou_multiple", "mean_iou", "max_iou")


def build_solver(config, num_aug=NUM_AUG, feature_size=FEATURE_SIZE, output_size=IMG_SIZE):
    """Optimizer + Superresolution as sweep_script.py:88-93 builds them from `wandb.config`."""
    c = dict(HYPERPARAMETERS_DEFAULT, **config)
    optimizer_obj = Optimizer(optimizer=c["optimizer"], learning_rate=c["learning_rate"], epsilon=c["epsilon"], beta_1=c["beta_1"], beta_2=c["beta_2"],
                              amsgrad=c["amsgrad"], initial_accumulator_value=c["initial_accumulator_value"], momentum=c["momentum"],
                              nesterov=c["nesterov"], lr_scheduler=c["lr_scheduler"], decay_steps=c["decay_steps"], decay_rate=c["decay_rate"])
    return Superresolution(lambda_df=c["lambda_df"], lambda_tv=c["lambda_tv"], lambda_L2=c["lambda_L2"], lambda_L1=c["lambda_L1"], num_iter=c["num_iter"],
                           num_aug=num_aug, optimizer=optimizer_obj, use_BTV=c["use_BTV"], copy_dropout=c["copy_dropout"],
                           feature_size=feature_size, output_size=output_size)


def run_grid(configs, precomputed_dir, true_mask_dir, standard_mask_dir=None, num_aug=NUM_AUG, num_samples=NUM_SAMPLES, class_id=CLASS_ID,
             th_factor=TH_FACTOR, batch=16, img_size=IMG_SIZE, feature_size=FEATURE_SIZE, verbose=True):
    """Every configuration of `configs` (dicts overriding HYPERPARAMETERS_DEFAULT) over the directory, `batch` images at a time.
    Returns a list of {wandb.log key: average} dicts, one per configuration."""
    torch = _lib._torch()
    np.random.seed(SEED)
    solvers = [build_solver(c, num_aug, feature_size, img_size) for c in configs]
    if any(s.copy_dropout for s in solvers):
        raise NotImplementedError("run_grid shares one launch between the points: per-point copy dropout masks are not supported, use run_point")
    P = len(solvers)
    paths = list_precomputed_data_paths(precomputed_dir, sort=True)
    paths = paths if num_samples is None else paths[:num_samples]
    ious = [{k: [] for k in LOG_KEYS} for _ in range(P)]
    n_done = 0          # valid images so far: the step offset of a point is (solves per image) * n_done * num_iter of that point

    def flush(pending):
        nonlocal n_done
        if not pending:
            return
        cls = torch.stack([p[1][..., 0] for p in pending]).contiguous()
        has_max = pending[0][2] is not None
        stacks = torch.cat([cls, torch.stack([p[2][..., 0] for p in pending]).contiguous()]) if has_max else cls     # class stacks, then max stacks
        B = len(pending)
        ang = np.stack([np.asarray(p[3], np.float32) for p in pending]); shf = np.stack([np.asarray(p[4], np.float32) for p in pending])
        if has_max:
            ang, shf = np.concatenate([ang, ang]), np.concatenate([shf, shf])
        per_image = 2 if has_max else 1
        plist, index = [], []
        for pi, s in enumerate(solvers):
            n_it = int(s.num_iter)
            for j in range(B):
                plist.append(s._solve_params(per_image * (n_done + j) * n_it)); index.append(j)
                if has_max:
                    plist.append(s._solve_params((per_image * (n_done + j) + 1) * n_it)); index.append(B + j)
        x = _lib.solve_sweep(stacks, ang, shf, plist, index).reshape(P, B, per_image, *img_size)
        true = np.stack([load_image(os.path.join(true_mask_dir, f"{p[5]}.png"), image_size=img_size, normalize=False, is_png=True,
                                    resize_method="nearest")[..., 0] for p in pending])
        true_d = torch.from_numpy(true.astype(np.int32)).cuda()
        # max / mean back-projection and the standard masks do not depend on the configuration
        fixed = {}
        for kind in ("max", "mean"):
            xc = solvers[0].backproject_batched(cls, ang[:B], shf[:B], kind)
            xm = solvers[0].backproject_batched(stacks[B:], ang[:B], shf[:B], kind) if has_max else None
            fixed[kind] = iou_table(true_d, _threshold_batched(xc, class_id, th_factor, th_mask=xm), class_id)[0]
        std = None
        if standard_mask_dir is not None:
            sm = np.stack([load_image(os.path.join(standard_mask_dir, f"{p[5]}.png"), image_size=img_size, normalize=False, is_png=True,
                                      resize_method="nearest")[..., 0] for p in pending])
            std = iou_table(true_d, torch.from_numpy(sm.astype(np.int32)).cuda(), class_id)
        for pi in range(P):
            xc = x[pi, :, 0].contiguous()
            xm = x[pi, :, 1].contiguous() if has_max else None
            a1, a2 = iou_table(true_d, _threshold_batched(xc, class_id, th_factor, th_mask=xm), class_id)
            ious[pi]["aug_iou_single"] += list(a1); ious[pi]["aug_iou_multiple"] += list(a2)
            ious[pi]["max_iou"] += list(fixed["max"]); ious[pi]["mean_iou"] += list(fixed["mean"])
            if std is not None:
                ious[pi]["standard_iou_single"] += list(std[0]); ious[pi]["standard_iou_multiple"] += list(std[1])
        n_done += B

    pending = []
    for path in paths:
        try:
            item = load_SR_data(path, num_aug=num_aug, global_normalize=True)
        except Exception:
            if verbose:
                print(f"File: {path} is invalid, skipping...")
            continue
        if pending and ((pending[0][2] is None) != (item[1] is None)):
            flush(pending); pending = []
        pending.append((path,) + tuple(item))
        if len(pending) >= batch:
            flush(pending); pending = []
    flush(pending)
    out = [{k: (float(np.mean(v)) if len(v) else float("nan")) for k, v in d.items()} for d in ious]
    if verbose:
        for c, o in zip(configs, out):
            print(json.dumps({"config": c, **o}))
    return out


def run_point(config, precomputed_dir, true_mask_dir, standard_mask_dir=None, **kw):
    """The reference's main() for one `wandb.config`: returns the dict it logs (:163-170)."""
    return run_grid([config], precomputed_dir, true_mask_dir, standard_mask_dir, **kw)[0]


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--data-dir", default=os.path.join(os.getcwd(), "data"))
    ap.add_argument("--grid", default=None, help="JSON file with a list of configuration overrides (default: the single default point)")
    ap.add_argument("--num-aug", type=int, default=NUM_AUG)
    ap.add_argument("--num-samples", type=int, default=NUM_SAMPLES)
    ap.add_argument("--class-id", type=int, default=CLASS_ID)
    ap.add_argument("--mode", default=MODE)
    ap.add_argument("--backbone", default=MODEL_BACKBONE)
    ap.add_argument("--validation", action="store_true")
    ap.add_argument("--th-factor", type=float, default=TH_FACTOR)
    ap.add_argument("--batch", type=int, default=16, help="images per launch (each carries every grid point)")
    a = ap.parse_args(argv)
    val = "_validation" if a.validation else ""
    pascal_root = os.path.join(a.data_dir, "dataset_root", "VOCdevkit", "VOC2012")
    superres_root = os.path.join(a.data_dir, "superres_root")
    precomputed = os.path.join(superres_root, "augmented_copies", f"{a.backbone}_{a.mode}_{a.class_id}_{a.num_aug}{val}")
    standard = os.path.join(superres_root, "standard_output", f"{a.backbone}_{a.class_id}{val}")
    grid = json.load(open(a.grid)) if a.grid else [{}]
    return run_grid(grid, precomputed, os.path.join(pascal_root, "SegmentationClassAug"), standard if os.path.isdir(standard) else None,
                    num_aug=a.num_aug, num_samples=a.num_samples, class_id=a.class_id, th_factor=a.th_factor, batch=a.batch)


if __name__ == "__main__":
    main()
