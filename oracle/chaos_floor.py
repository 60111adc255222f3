import time, numpy as np, sys
sys.path.insert(0, "/root/repo")
from oracle import oracle as O
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
copies, ang, sh = make_augmented_copies(1, 100, value=1.0)
copies = copies[0].numpy(); ang = ang[0]; sh = sh[0]
P = O.SolveParams(num_iter=300)
ti = [1, 2, 5, 10, 20, 50, 100, 150, 200, 250, 300]
t=time.time(); x, loss, tr = O.augmented_superresolution(copies, ang, sh, P, trace_iters=ti); print("base", time.time()-t, loss)
np.save("/tmp/base_trace.npy", tr)
t=time.time(); x2, loss2, tr2 = O.augmented_superresolution(copies, ang, sh, P, trace_iters=ti, variant="_fma"); print("fma", time.time()-t, loss2)
np.save("/tmp/fma_trace.npy", tr2)
for i, it in enumerate(ti):
    d = np.abs(tr[i]-tr2[i])
    print(it, "maxabs", d.max(), "mean", d.mean(), "n>1e-4", (d>1e-4).sum(), "n>1e-5", (d>1e-5).sum())
th1 = O.threshold_image(tr[-1], 1, 0.65); th2 = O.threshold_image(tr2[-1], 1, 0.65)
print("mask agreement", (th1==th2).mean())
