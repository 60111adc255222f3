"""Small solves + aux calls for compute-sanitizer (memcheck / racecheck / synccheck / initcheck).
   compute-sanitizer --tool memcheck python scripts/dev_sanitize.py
   (On this GPU pool compute-sanitizer is closed -- "runs under it have left GPUs needing a reset" -- so no output is kept; bad
   accesses are guarded by the __trap() bounds checks of the box computations and by bit-exact comparison with the oracle.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deeplabv3plus_augmented_superresolution_b200 import _lib as A
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts import augmentation_utils as AU, superres_utils as SU
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.optimizer import Optimizer
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution

# x4 (tuned kernels): small angles (60/76-row box) and any angle (big box), ragged sizes, two tile heights of K2
for (hw, amax, B) in (((32, 32), 0.15, 2), ((24, 40), 2.5, 1), ((48, 48), 0.15, 3)):
    copies, ang, sh = make_augmented_copies(B, 7, hw, (4 * hw[0], 4 * hw[1]), amax, 25, seed=5, device="cuda")
    x, loss = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=3), want_loss=True)
    torch.cuda.synchronize()
    print("solve x4", hw, amax, float(x.sum()), float(loss.sum()))
# another even ratio (literal kernels)
copies, ang, sh = make_augmented_copies(1, 4, (16, 16), (128, 128), 0.3, 10, seed=6, device="cuda")
x = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=2), output_size=(128, 128))
torch.cuda.synchronize()
print("solve x8", float(x.sum()))
# aux kernels
img = torch.rand((40, 56, 3), device="cuda")
w = AU.warp_copies(img, np.array([0.0, 0.3], np.float32), np.array([[0, 0], [5.5, -3]], np.float32), "bilinear")
logits = torch.randn((3, 24, 40, 21), device="cuda")
for mode in ("argmax", "slice", "slice_max"):
    c, m = AU.extract_opm(logits, 8, mode)
sr = Superresolution(1, 0.3, 0.7, 0, optimizer=Optimizer(), feature_size=(16, 16), output_size=(64, 64))
copies, ang, sh = make_augmented_copies(2, 5, (16, 16), (64, 64), 0.3, 10, seed=7, device="cuda")
mx = sr.backproject_batched(copies, ang, sh, "max")
th = SU.threshold_image(mx[0].cpu().numpy()[..., None], 8, th_factor=0.2)
n = SU._normalize_stack_device(copies[0, :, :, :, None].contiguous())
torch.cuda.synchronize()
print("aux ok", float(w.sum()), float(mx.sum()), int(th.sum()), float(n.sum()))
