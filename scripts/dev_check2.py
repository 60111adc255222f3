import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as O
from deeplabv3plus_augmented_superresolution_b200 import _lib as A
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
h = 32
copies, ang, sh = make_augmented_copies(1, 6, (h, h), (4*h, 4*h), 0.15, 80, 1234, 1.0, device="cuda")
cp = copies.cpu().numpy()[0]
iters = list(range(1, 13))
xo, lo, tr = O.augmented_superresolution(cp, ang[0], sh[0], O.SolveParams(num_iter=12), output_size=(4*h, 4*h), trace_iters=iters)
for n in iters:
    x = A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=n))[0].cpu().numpy()
    d = np.abs(x - tr[n-1]); print(n, "identical", np.array_equal(x, tr[n-1]), "maxabs", d.max(), "ndiff", (x != tr[n-1]).sum(), flush=True)
    if not np.array_equal(x, tr[n-1]):
        # gradient at the oracle's previous x
        xp = torch.from_numpy(tr[n-2]).cuda()[None].contiguous()
        r, g, l = A.loss_grad_batched(xp, copies, ang, sh, A.SolveParams())
        lo2, go, ro = O.loss_and_grad(tr[n-2], cp, ang[0], sh[0], O.SolveParams(), want_resid=True)
        print("  grad at prev x identical:", np.array_equal(g[0].cpu().numpy(), go), "resid:", np.array_equal(r[0].cpu().numpy(), ro))
        idx = np.argwhere(x != tr[n-1])[:5]
        for (yy, xx) in idx:
            print("   px", yy, xx, "gpu", x[yy, xx], "orc", tr[n-1][yy, xx], "prev", tr[n-2][yy, xx], "g", go[yy, xx])
        break
