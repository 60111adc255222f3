"""One call of each auxiliary entry point at the reference shapes (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts import augmentation_utils as AU, superres_utils as SU
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.optimizer import Optimizer
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution
g = torch.Generator(device="cuda").manual_seed(0)
img = torch.rand((512, 512, 3), device="cuda", generator=g)
np.random.seed(1234)
ang, sh = AU._draw(100, 0.15, 80)
logits = torch.randn((100, 128, 128, 21), device="cuda", generator=g)
copies, a2, s2 = make_augmented_copies(4, 100, device="cuda")
sr = Superresolution(1, 0.3, 0.7, 0, optimizer=Optimizer(), feature_size=(128, 128))
for rep in range(2):
    AU.warp_copies(img, ang, sh, "bilinear"); AU.warp_copies(img, ang, sh, "nearest")
    for m in ("argmax", "slice", "slice_max"):
        AU.extract_opm(logits, 8, m)
    SU._normalize_stack_device(copies[0, :, :, :, None].contiguous())
    sr.backproject_batched(copies, a2, s2, "max"); sr.backproject_batched(copies, a2, s2, "mean")
    SU.threshold_image(torch.rand((512, 512, 1), device="cuda"), 8, th_factor=0.65)
torch.cuda.synchronize()
print("ok")
