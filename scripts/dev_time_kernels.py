import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplabv3plus_augmented_superresolution_b200 import _lib as A
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
L = A.lib()
copies, ang, sh = make_augmented_copies(B, 100, device="cuda")
A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=5)); torch.cuda.synchronize()
L.asr_profile_enable(1)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=iters)); e1.record(); torch.cuda.synchronize()
ms = (C.c_double * 2)(); cnt = (C.c_longlong * 2)()
L.asr_profile_read(ms, cnt); L.asr_profile_enable(0)
tot = e0.elapsed_time(e1)
print(f"B={B} iters={iters}: total {tot:.1f} ms -> {tot/iters/B*1e3:.1f} us/image-iter | K1 {ms[0]/cnt[0]/B*1e3:.1f} us/image | K2 {ms[1]/cnt[1]/B*1e3:.1f} us/image | images/s at 300 it: {B/(tot/iters*300)*1e3:.1f}")
