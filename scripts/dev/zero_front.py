"""How long does the exact-zero background of x survive the solve?  (exact-zero skipping.)
Uses the CPU oracle (test infrastructure) on the config-1 synthetic image; output kept in profiles/r02_zero_front.txt."""
import sys, numpy as np
sys.path.insert(0,'"" or __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))')
from oracle import oracle as O
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
O.use_all_cores()
copies, ang, sh = make_augmented_copies(1, 100, (128,128), (512,512), 0.15, 80, seed=1234, value=1.0)
c = copies[0].numpy()
print("LR nonzero fraction", (c!=0).mean())
for it in (1,2,5,10,20,40):
    x,_ = O.augmented_superresolution(c, ang[0], sh[0], O.SolveParams(num_iter=it), output_size=(512,512))
    print(it, "x exact-zero fraction %.4f" % (x==0).mean(), flush=True)
