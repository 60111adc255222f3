"""Development check (GPU box): CUDA solve vs CPU oracle, bitwise, plus a first timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as O
from deeplabv3plus_augmented_superresolution_b200 import _lib as A
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies

def cmp(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    print(f"{name}: bit-identical={np.array_equal(a, b)} maxabs={d.max():.3e} n_diff={(a != b).sum()} / {a.size}", flush=True)

def case(B, N, h, iters, value=1.0, angle_max=0.15, shift_max=80, seed=1234, **kw):
    print(f"--- case B={B} N={N} h={h} iters={iters} angle_max={angle_max} shift_max={shift_max} {kw}", flush=True)
    copies, ang, sh = make_augmented_copies(B, N, (h, h), (4*h, 4*h), angle_max, shift_max, seed, value, device="cuda")
    P = A.SolveParams(num_iter=iters, **kw); PO = O.SolveParams(num_iter=iters, **{k: v for k, v in kw.items()})
    cp = copies.cpu().numpy()
    x0 = torch.stack([torch.from_numpy(O.resize_bilinear(cp[b, :1, :, :, None], (4*h, 4*h))[0, :, :, 0]) for b in range(B)]).cuda()
    r, g, l = A.loss_grad_batched(x0, copies, ang, sh, P)
    for b in range(B):
        lo, go, ro = O.loss_and_grad(x0[b].cpu().numpy(), cp[b], ang[b], sh[b], PO, want_resid=True)
        cmp(f" resid[{b}]", r[b].cpu().numpy(), ro); cmp(f" grad[{b}]", g[b].cpu().numpy(), go)
        print("  loss", float(l[b]), lo)
    x, loss = A.solve_batched(copies, ang, sh, P, want_loss=True)
    torch.cuda.synchronize()
    for b in range(B):
        xo, lo = O.augmented_superresolution(cp[b], ang[b], sh[b], PO, output_size=(4*h, 4*h))
        cmp(f" x[{b}] after {iters}", x[b].cpu().numpy(), xo[..., 0]); print("  loss", float(loss[b]), lo)

case(2, 6, 32, 20)
case(1, 5, 32, 10, angle_max=3.1, shift_max=30, seed=7)
case(1, 4, 16, 10, angle_max=0.5, shift_max=70, seed=3, value=8.0)
case(1, 100, 128, 3)
# timing
for B in (1, 8, 32):
    copies, ang, sh = make_augmented_copies(B, 100, device="cuda")
    P = A.SolveParams(num_iter=300)
    A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=5)); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); x = A.solve_batched(copies, ang, sh, P); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"B={B}: {ms:.1f} ms  -> {B/ms*1e3:.1f} images/s  ({ms/300/B*1e3:.1f} us per image-iteration)", flush=True)
