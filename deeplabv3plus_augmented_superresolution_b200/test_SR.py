"""B200 counterpart of the reference's test_SR.py (:1-104): the single-image demo.

One image -> NUM_AUG rotated/translated copies (create_augmented_copies) -> upstream model.predict -> OPM extraction ->
augmented / max / mean super-resolution -> thresholded masks -> IoU against the ground truth.  Same constants as the
reference.  The DeepLabV3+ forward is an upstream producer outside this repo: pass any object with
`predict(images, batch_size) -> [n,128,128,21] logits` as `model`; `--synthetic` (and the test-suite) use the analytic
stand-in of synthetic.py because neither the weights nor TensorFlow exist here.

    python -m deeplabv3plus_augmented_superresolution_b200.test_SR --synthetic
"""
import argparse
import os
import tempfile

import numpy as np

from .superresolution_scripts.augmentation_utils import compute_augmented_feature_maps
from .superresolution_scripts.optimizer import Optimizer
from .superresolution_scripts.superresolution import Superresolution
from .superresolution_scripts.superres_utils import compute_SR
from .utils import compute_IoU, load_image

SEED = 1234

IMG_SIZE = (512, 512)
FEATURE_SIZE = (128, 128)
BATCH_SIZE = 16
CLASS_ID = 8
MODE = "argmax"

NUM_AUG = 100
ANGLE_MAX = 0.15
SHIFT_MAX = 80

OPTIMIZER = "adam"
LEARNING_RATE = 1e-3
AMSGRAD = True
LR_SCHEDULER = True
DECAY_STEPS = 60
DECAY_RATE = 0.3

LAMBDA_DF = 1.0
LAMBDA_TV = 0.3
LAMBDA_L2 = 0.7
LAMBDA_L1 = 0.0
NUM_ITER = 300
TH_FACTOR = 0.2


def run(img_path, gt_path, model, sr_output_dir, mode=MODE, num_aug=NUM_AUG, num_iter=NUM_ITER, save=True, verbose=True):
    """The body of the reference's main() (:57-100).  Returns {"aug": iou, "max": iou, "mean": iou} and the three masks."""
    np.random.seed(SEED)
    optimizer_obj = Optimizer(optimizer=OPTIMIZER, learning_rate=LEARNING_RATE, amsgrad=AMSGRAD,
                              lr_scheduler=LR_SCHEDULER, decay_steps=DECAY_STEPS, decay_rate=DECAY_RATE)
    superresolution_obj = Superresolution(lambda_df=LAMBDA_DF, lambda_tv=LAMBDA_TV, lambda_L2=LAMBDA_L2, lambda_L1=LAMBDA_L1,
                                          num_iter=num_iter, num_aug=num_aug, optimizer=optimizer_obj, feature_size=FEATURE_SIZE)

    class_masks, max_masks, angles, shifts, filename = compute_augmented_feature_maps(
        img_path, model, filter_class_id=CLASS_ID, mode=mode, num_aug=num_aug, angle_max=ANGLE_MAX,
        shift_max=SHIFT_MAX, image_size=IMG_SIZE, batch_size=BATCH_SIZE)

    masks = {}
    for sr_type in ("aug", "max", "mean"):
        masks[sr_type] = compute_SR(superresolution_obj, class_masks, angles, shifts, filename, max_masks=max_masks, SR_type=sr_type,
                                    save_final_output=save, class_id=CLASS_ID, dest_folder=sr_output_dir, th_factor=TH_FACTOR)

    gt_mask = load_image(gt_path, image_size=IMG_SIZE, normalize=False, is_png=True, resize_method="nearest")
    ious = {k: compute_IoU(gt_mask, m, img_size=IMG_SIZE, class_id=CLASS_ID) for k, m in masks.items()}
    if verbose:
        print(f"Aug. SR ({mode} OPM) IoU: {ious['aug']}, Max SR IoU: {ious['max']}, Mean SR IoU: {ious['mean']}")
    return ious, masks


def write_synthetic_inputs(folder):
    """test_cat.jpg / test_cat_gt.png stand-ins (synthetic.make_test_image) as files, the way the reference reads them."""
    from PIL import Image
    from .synthetic import make_test_image
    img, gt = make_test_image(IMG_SIZE, CLASS_ID)
    os.makedirs(folder, exist_ok=True)
    img_path, gt_path = os.path.join(folder, "test_shape.png"), os.path.join(folder, "test_shape_gt.png")
    Image.fromarray((img * 255.0 + 0.5).astype(np.uint8), mode="RGB").save(img_path)
    Image.fromarray(gt[..., 0].astype(np.uint8), mode="L").save(gt_path)
    return img_path, gt_path


def main(argv=None, model=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--image", default=os.path.join(os.getcwd(), "test_images", "test_cat.jpg"))
    ap.add_argument("--gt", default=os.path.join(os.getcwd(), "test_images", "test_cat_gt.png"))
    ap.add_argument("--out", default=os.path.join(os.getcwd(), "test_images", "SR_output"))
    ap.add_argument("--mode", default=MODE)
    ap.add_argument("--synthetic", action="store_true", help="analytic image, ground truth and upstream model (no DeepLab weights needed)")
    a = ap.parse_args(argv)
    if a.synthetic:
        from .synthetic import SyntheticSegmenter
        tmp = tempfile.mkdtemp(prefix="asr_test_sr_")
        a.image, a.gt = write_synthetic_inputs(tmp)
        a.out = os.path.join(tmp, "SR_output")
        model = SyntheticSegmenter(classes=21, class_id=CLASS_ID)
    if model is None:
        raise SystemExit("test_SR needs the upstream DeepLabV3+ model (model.py in the reference, outside this repo): "
                         "call main(model=...) with an object exposing predict(images, batch_size), or pass --synthetic")
    return run(a.image, a.gt, model, a.out, mode=a.mode)


if __name__ == "__main__":
    main()
