"""Small solve + aux calls for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplabv3plus_augmented_superresolution_b200 import _lib as A
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
copies, ang, sh = make_augmented_copies(2, 7, (32, 32), (128, 128), 0.4, 25, seed=5, device="cuda")
x = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=4))
torch.cuda.synchronize()
print("solve ok", float(x.sum()))
