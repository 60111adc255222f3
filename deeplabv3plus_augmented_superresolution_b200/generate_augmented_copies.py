"""B200 counterpart of the reference's generate_augmented_copies.py (:1-103): writes the hdf5 augmented-copies files that
SR_single_class / sweep_script read.

Same command line, same directory naming (`<backbone>_<mode>_<class>_<num_aug>[_validation]`), same global NumPy seed, same
image selection (the first `num_samples` class-filtered images of the sorted list).  Per image: create_augmented_copies (warp on
the device) -> `model.predict` -> OPM extraction on the device -> one hdf5 file (augmentation_utils.py:117-136 layout).
The DeepLabV3+ forward is an upstream producer outside this repo: `main(argv, model=...)` takes any object with
`predict(images, batch_size) -> [n,128,128,21] logits`; `--synthetic-model` substitutes the analytic stand-in of synthetic.py.
"""
import argparse
import os

import numpy as np

from .superresolution_scripts.augmentation_utils import compute_augmented_feature_maps
from .superresolution_scripts.superres_utils import filter_images_by_class, get_img_paths

SEED = 1234
IMG_SIZE = (512, 512)
BATCH_SIZE = 16


def build_parser():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--num_aug", help="Number of augmented copies created for each image", action="store", type=int, default=100)
    parser.add_argument("--num_samples", help="Number of samples taken from the dataset", action="store", type=int, default=500)
    parser.add_argument("--mode", help="Whether to operate in slicing, slicing variation or argmax mode", action="store", type=str,
                        choices=["slice_max", "slice", "argmax"], default="argmax")
    parser.add_argument("--angle_max", help="Max angle value (in radians) used for rotations", action="store", type=float, default=0.3)
    parser.add_argument("--shift_max", help="Max shift value used for traslations", action="store", type=int, default=30)
    parser.add_argument("--backbone", help="Either mobilenet or xception, specifies the type of backbone to use", action="store", type=str,
                        choices=["mobilenet", "xception"], default="xception")
    parser.add_argument("--use_validation", help="Create data from validation set", action="store_true")
    parser.add_argument("--class_id", help="class_id for image filtering", action="store", type=int, default=8, choices=range(21), required=True)
    parser.add_argument("--data_dir", help="root of the data tree (the reference uses ./data)", default=os.path.join(os.getcwd(), "data"))
    parser.add_argument("--synthetic-model", action="store_true", help="use the analytic stand-in for DeepLabV3+ (no weights needed)")
    return parser


def run(args, model, image_size=IMG_SIZE, verbose=True):
    """The body of the reference's main() (:67-99).  Returns the list of hdf5 paths written."""
    np.random.seed(SEED)
    pascal_root = os.path.join(args.data_dir, "dataset_root", "VOCdevkit", "VOC2012")
    imgs_path = os.path.join(pascal_root, "JPEGImages")
    out_dir = os.path.join(args.data_dir, "superres_root", "augmented_copies",
                           f"{args.backbone}_{args.mode}_{args.class_id}_{args.num_aug}{'_validation' if args.use_validation else ''}")
    image_list_path = os.path.join(args.data_dir, "augmented_file_lists", f"{'valaug' if args.use_validation else 'trainaug'}.txt")
    image_paths = get_img_paths(image_list_path, imgs_path, is_png=False, sort=True)
    images_paths_filtered = filter_images_by_class(image_paths, filter_class_id=args.class_id, num_images=args.num_samples, image_size=image_size)
    if verbose:
        print(f"Valid images: {len(images_paths_filtered)} (Initial: {len(image_paths)})")
        print("Generating augmented copies...")
    written = []
    for image_path in images_paths_filtered:
        *_, name = compute_augmented_feature_maps(image_path, model, mode=args.mode, filter_class_id=args.class_id, num_aug=args.num_aug,
                                                  angle_max=args.angle_max, shift_max=args.shift_max, image_size=image_size,
                                                  batch_size=BATCH_SIZE, dest_folder=out_dir)
        written.append(os.path.join(out_dir, f"{name}.hdf5"))
    return written


def main(argv=None, model=None):
    args = build_parser().parse_args(argv)
    if args.synthetic_model:
        from .synthetic import SyntheticSegmenter
        model = SyntheticSegmenter(classes=21, class_id=args.class_id)
    if model is None:
        raise SystemExit("generate_augmented_copies needs the upstream DeepLabV3+ model (model.py in the reference, outside this repo): "
                         "call main(argv, model=...) with an object exposing predict(images, batch_size), or pass --synthetic-model")
    return run(args, model)


if __name__ == "__main__":
    main()
