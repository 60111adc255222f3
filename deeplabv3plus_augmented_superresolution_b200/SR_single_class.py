"""B200 counterpart of the reference's SR_single_class.py (:1-145): the Augmented Super-Resolution batch run over a
directory of precomputed hdf5 augmented-copies files (written by generate_augmented_copies.py or by
compute_augmented_feature_maps(dest_folder=...)).

Same constants, same hyper-parameters, same three printed lines.  The only difference is the schedule: instead of one image
per iteration of a Python loop, `--batch` images go through every kernel launch (batch_runner.run_files); the numbers are
those the sequential loop produces, including the shared optimizer's running step counter.

    python -m deeplabv3plus_augmented_superresolution_b200.SR_single_class --data-dir data [--num-aug 10] [--mode argmax]
"""
import argparse
import os

import numpy as np

from .batch_runner import iou_table, run_files
from .superresolution_scripts.optimizer import Optimizer
from .superresolution_scripts.superresolution import Superresolution
from .superresolution_scripts.superres_utils import list_precomputed_data_paths
from .utils import load_image

SEED = 1234

IMG_SIZE = (512, 512)
FEATURE_SIZE = (128, 128)
NUM_AUG = 10
CLASS_ID = 8
NUM_SAMPLES = 500
MODE = "argmax"
MODEL_BACKBONE = "xception"
USE_VALIDATION = True
TH_FACTOR = 0.65

HYPERPARAMETERS = {
    "lambda_df": 1,
    "lambda_tv": 0.3,
    "lambda_L2": 0.7,
    "lambda_L1": 0.0,
    "num_iter": 300,
    "optimizer": "adam",
    "learning_rate": 1e-3,
    "amsgrad": True,
    "lr_scheduler": True,
    "decay_steps": 60,
    "decay_rate": 0.3,
}


def build_solver(num_aug=NUM_AUG, feature_size=FEATURE_SIZE, output_size=IMG_SIZE, **overrides):
    """Optimizer + Superresolution exactly as SR_single_class.py:66-71 builds them."""
    hp = dict(HYPERPARAMETERS, **overrides)
    optimizer_obj = Optimizer(optimizer=hp["optimizer"], learning_rate=hp["learning_rate"], amsgrad=hp["amsgrad"],
                              lr_scheduler=hp["lr_scheduler"], decay_steps=hp["decay_steps"], decay_rate=hp["decay_rate"])
    return Superresolution(lambda_df=hp["lambda_df"], lambda_tv=hp["lambda_tv"], lambda_L2=hp["lambda_L2"], lambda_L1=hp["lambda_L1"],
                           num_iter=hp["num_iter"], num_aug=num_aug, optimizer=optimizer_obj, feature_size=feature_size,
                           output_size=output_size)


def run(precomputed_dir, true_mask_dir, standard_mask_dir=None, num_aug=NUM_AUG, num_samples=NUM_SAMPLES, class_id=CLASS_ID,
        th_factor=TH_FACTOR, batch=64, img_size=IMG_SIZE, feature_size=FEATURE_SIZE, verbose=True, **hyper):
    """The body of the reference's main() (:73-141).  Returns the six averages as a dict."""
    import torch
    np.random.seed(SEED)
    sr = build_solver(num_aug=num_aug, feature_size=feature_size, output_size=img_size, **hyper)
    path_list = list_precomputed_data_paths(precomputed_dir, sort=True)
    paths = path_list if num_samples is None else path_list[:num_samples]

    ious = {k: [] for k in ("standard_single", "standard_multiple", "aug_single", "aug_multiple", "max", "mean")}

    def skip(p):
        if verbose:
            print(f"File: {p} is invalid, skipping...")

    for res in run_files(sr, paths, num_aug=num_aug, class_id=class_id, th_factor=th_factor, batch=batch, on_skip=skip):
        if not res.filenames:
            continue
        true = np.stack([load_image(os.path.join(true_mask_dir, f"{f}.png"), image_size=img_size, normalize=False, is_png=True,
                                    resize_method="nearest")[..., 0] for f in res.filenames])
        true_d = torch.from_numpy(true.astype(np.int32)).cuda()
        if standard_mask_dir is not None:
            std = np.stack([load_image(os.path.join(standard_mask_dir, f"{f}.png"), image_size=img_size, normalize=False, is_png=True,
                                       resize_method="nearest")[..., 0] for f in res.filenames])
            s1, s2 = iou_table(true_d, torch.from_numpy(std.astype(np.int32)).cuda(), class_id)
            ious["standard_single"] += list(s1); ious["standard_multiple"] += list(s2)
        a1, a2 = iou_table(true_d, res.aug, class_id)
        ious["aug_single"] += list(a1); ious["aug_multiple"] += list(a2)
        ious["max"] += list(iou_table(true_d, res.max, class_id)[0])
        ious["mean"] += list(iou_table(true_d, res.mean, class_id)[0])

    avg = {k: (float(np.mean(v)) if len(v) else float("nan")) for k, v in ious.items()}   # np.mean, NaNs propagate as in the reference
    avg["images"] = len(ious["aug_single"])
    if verbose:
        print(f"Avg. Standard IoUs (No bg): {avg['standard_single']},  Avg. Augmented SR IoUs (No bg): {avg['aug_single']}")
        print(f"Avg. Standard IoUs (with bg): {avg['standard_multiple']},  Avg. Augmented SR IoUs (with bg): {avg['aug_multiple']}")
        print(f"Avg. Max SR IoUs: {avg['max']}, Avg. Mean SR IoUs: {avg['mean']}")
    return avg


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--data-dir", default=os.path.join(os.getcwd(), "data"))
    ap.add_argument("--num-aug", type=int, default=NUM_AUG)
    ap.add_argument("--num-samples", type=int, default=NUM_SAMPLES)
    ap.add_argument("--class-id", type=int, default=CLASS_ID)
    ap.add_argument("--mode", default=MODE)
    ap.add_argument("--backbone", default=MODEL_BACKBONE)
    ap.add_argument("--no-validation", action="store_true")
    ap.add_argument("--th-factor", type=float, default=TH_FACTOR)
    ap.add_argument("--batch", type=int, default=64, help="images per kernel launch")
    a = ap.parse_args(argv)
    val = "" if a.no_validation else "_validation"
    pascal_root = os.path.join(a.data_dir, "dataset_root", "VOCdevkit", "VOC2012")
    superres_root = os.path.join(a.data_dir, "superres_root")
    precomputed = os.path.join(superres_root, "augmented_copies", f"{a.backbone}_{a.mode}_{a.class_id}_{a.num_aug}{val}")
    standard = os.path.join(superres_root, "standard_output", f"{a.backbone}_{a.class_id}{val}")
    return run(precomputed, os.path.join(pascal_root, "SegmentationClassAug"), standard, num_aug=a.num_aug,
               num_samples=a.num_samples, class_id=a.class_id, th_factor=a.th_factor, batch=a.batch)


if __name__ == "__main__":
    main()
