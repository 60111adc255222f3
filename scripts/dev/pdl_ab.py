"""Whole-solve time with and without programmatic dependent launch (run once with ASR_PDL=1, once with ASR_PDL=0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from deeplabv3plus_augmented_superresolution_b200 import _lib as A
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies

for B, iters in ((1, 300), (2, 300), (3, 200), (4, 150), (6, 100), (8, 100), (64, 20)):
    copies, ang, sh = make_augmented_copies(B, 100, device="cuda")
    A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=5)); torch.cuda.synchronize()
    best = 1e9
    for rep in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); x = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=iters)); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"lib={os.path.basename(A.LIB_PATH)} ASR_PDL={os.environ.get('ASR_PDL', '3')} B={B:3d} iters={iters}: {best:8.2f} ms  = {best / iters * 1e3:7.1f} us/iteration  checksum {float(x.double().sum()):.6f}")
