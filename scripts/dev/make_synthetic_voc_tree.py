"""Builds a tiny synthetic data/ tree in the layout SR_single_class.py expects (SR_single_class.py:34-46): hdf5 augmented-copies
files, ground-truth PNGs and "standard output" PNGs, so that the entry point can be run as a command line:
    python scripts/dev/make_synthetic_voc_tree.py /tmp/asr_data 6 10
    python -m deeplabv3plus_augmented_superresolution_b200.SR_single_class --data-dir /tmp/asr_data --num-aug 10 --batch 4"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from PIL import Image
from deeplabv3plus_augmented_superresolution_b200 import hdf5_lite
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies

root, n_img, n_aug = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
voc = os.path.join(root, "dataset_root", "VOCdevkit", "VOC2012")
gt_dir = os.path.join(voc, "SegmentationClassAug")
cop_dir = os.path.join(root, "superres_root", "augmented_copies", f"xception_argmax_8_{n_aug}_validation")
std_dir = os.path.join(root, "superres_root", "standard_output", "xception_8_validation")
for d in (gt_dir, cop_dir, std_dir):
    os.makedirs(d, exist_ok=True)
copies, ang, sh = make_augmented_copies(n_img, n_aug, (128, 128), (512, 512), 0.15, 80, seed=1234, value=8.0)
for b in range(n_img):
    name = f"2007_{b:06d}"
    f = hdf5_lite.File(os.path.join(cop_dir, name + ".hdf5"), "w")
    f.create_dataset("class_masks", data=[c[..., None] for c in copies[b].numpy()])
    f.create_dataset("angles", data=ang[b]); f.create_dataset("shifts", data=sh[b])
    f.attrs["filename"] = name; f.attrs["mode"] = "argmax"; f.attrs["angle_max"] = 0.15; f.attrs["shift_max"] = 80
    f.close()
    lr = copies[b, 0].numpy() > 0
    Image.fromarray((np.kron(lr, np.ones((4, 4))) * 8).astype(np.uint8), mode="L").save(os.path.join(gt_dir, name + ".png"))
    Image.fromarray((np.kron(lr, np.ones((4, 4))) * 8).astype(np.uint8), mode="L").save(os.path.join(std_dir, name + ".png"))
print("wrote", n_img, "images under", root)
