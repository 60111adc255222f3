"""Experiment: do two concurrent solves (two streams, half the images each) beat one solve of all images?
K1 is shared-memory bound and K2 fp32 bound, so co-resident CTAs of the two kernels could overlap."""
import sys, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplabv3plus_augmented_superresolution_b200 import _lib as A
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
copies, ang, sh = make_augmented_copies(B, 100, device="cuda")
P = A.SolveParams(num_iter=iters)
A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=3)); torch.cuda.synchronize()

def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)

t_one = timed(lambda: A.solve_batched(copies, ang, sh, P))
h = B // 2
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    def run(s, lo, hi, shift):
        with torch.cuda.stream(s):
            A.solve_batched(copies[lo:hi], ang[lo:hi], sh[lo:hi], P)
    # host threads so that the two launch sequences interleave
    ts = [threading.Thread(target=run, args=(s1, 0, h, 0)), threading.Thread(target=run, args=(s2, h, B, 1))]
    [t.start() for t in ts]; [t.join() for t in ts]
    cur.wait_stream(s1); cur.wait_stream(s2)
two()
t_two = timed(two)
print(f"B={B} iters={iters}: one solve {t_one:.1f} ms, two concurrent half solves {t_two:.1f} ms  (ratio {t_one / t_two:.3f})")
