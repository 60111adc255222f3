"""Short solve used under ncu: B images x 100 copies, a few iterations (kernel timings do not depend on the iteration)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplabv3plus_augmented_superresolution_b200 import _lib as A
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
copies, ang, sh = make_augmented_copies(B, 100, device="cuda")
x = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=iters))
torch.cuda.synchronize()
print("ok", float(x.sum()))
