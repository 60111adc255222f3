"""Generates tests/golden/*.npz with the CPU oracle (oracle/asr_oracle.c).  Run from the repo root:

    python tests/golden/make_golden.py [--canonical]

The reference itself (TensorFlow 2.7 + tensorflow-addons 0.15) cannot run in this image, so these are
outputs of the restated oracle, not of the reference: PARITY UNPINNED (see oracle/asr_oracle.h).
They pin the oracle across machines/compilers and give the GPU tests a fixed known answer.
Inputs are binary masks, stored bit-packed; outputs are stored as float32.
"""
import os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies

HERE = os.path.dirname(os.path.abspath(__file__))


def pack(copies, value):
    return np.packbits((copies != 0).astype(np.uint8).reshape(-1)), np.float32(value)


def solve_case(name, N, h, iters, value, angle_max, shift_max, seed, trace=(), **kw):
    copies, ang, sh = make_augmented_copies(1, N, (h, h), (4 * h, 4 * h), angle_max, shift_max, seed, value)
    copies = copies[0].numpy(); ang = ang[0]; sh = sh[0]
    P = O.SolveParams(num_iter=iters, **kw)
    t = time.time()
    res = O.augmented_superresolution(copies, ang, sh, P, output_size=(4 * h, 4 * h), trace_iters=list(trace))
    x, loss = res[0], res[1]
    tr = res[2] if trace else np.zeros((0, 4 * h, 4 * h), np.float32)
    x0 = O.resize_bilinear(copies[:1, :, :, None], (4 * h, 4 * h))[0, :, :, 0]
    l0, g0, r0 = O.loss_and_grad(x0, copies, ang, sh, P, want_resid=True)
    mx = O.backproject(copies, ang, sh, "max", (4 * h, 4 * h))
    mn = O.backproject(copies, ang, sh, "mean", (4 * h, 4 * h))
    bits, val = pack(copies, value)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), copies_bits=bits, value=val, shape=np.array(copies.shape),
                        angles=ang, shifts=sh, x=x[..., 0], loss=np.float32(loss), trace_iters=np.array(list(trace), np.int32),
                        trace=tr, grad0=g0, resid0_sum=np.float64(r0.astype(np.float64).sum()), loss0=np.float32(l0),
                        max_sr=mx[..., 0], mean_sr=mn[..., 0], params=np.array(repr(P)))
    print(f"{name}: {time.time() - t:.1f}s loss={loss}")


if __name__ == "__main__":
    solve_case("small_adam", N=8, h=32, iters=40, value=1.0, angle_max=0.15, shift_max=20, seed=11, trace=(1, 5, 20))
    solve_case("small_value8_bigangle", N=6, h=32, iters=25, value=8.0, angle_max=1.2, shift_max=30, seed=5,
               lambda_l1=0.05, amsgrad=False, step_offset=300)
    solve_case("small_sgd", N=5, h=16, iters=15, value=1.0, angle_max=0.3, shift_max=10, seed=2, optimizer="sgd",
               momentum=0.9, nesterov=True, learning_rate=1e-4, lr_scheduler=False)
    if "--canonical" in sys.argv:
        # config 1 of BASELINE.json: test_SR.py constants, N=100, 128^2 -> 512^2, 300 iterations
        solve_case("canonical_config1", N=100, h=128, iters=300, value=8.0, angle_max=0.15, shift_max=80, seed=1234,
                   trace=(1, 10, 100))
