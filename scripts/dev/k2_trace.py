"""Development aid: clock64 stamps of one K2 CTA's fill/gather hand-offs (build with -DASR_K2_TRACE).
usage: ASR_LIB=<trace build> python scripts/dev/k2_trace.py [B] -> gpurun_out/k2_trace.npz + a summary on stdout"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from deeplabv3plus_augmented_superresolution_b200 import _lib as A
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
copies, ang, sh = make_augmented_copies(B, 100, device="cuda")
A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=4))
torch.cuda.synchronize()
L = A.lib()
tr = np.zeros((16, 128, 4), np.int64); misc = np.zeros((16, 4), np.int64)
L.asr_debug_k2_trace.argtypes = [C.c_void_p, C.c_void_p]
assert L.asr_debug_k2_trace(tr.ctypes.data, misc.ctypes.data) == 0
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/k2_trace.npz", trace=tr, misc=misc)
t0 = misc[:, 0].min()
gw = 8
g = tr[:gw, :100]; f = tr[gw:12, :100]
wait = (g[:, :, 1] - g[:, :, 0]); work = (g[:, :, 2] - g[:, :, 1])
print("CTA lifetime clk", misc[:12, 3].max() - t0, "loop start", (misc[:gw, 1] - t0).mean(), "loop end", (misc[:gw, 2] - t0).mean(), "epilogue", (misc[:gw, 3] - misc[:gw, 2]).mean())
print("gather: wait-for-full per copy mean %.0f clk, gather per copy mean %.0f clk (per warp means: wait %s work %s)" % (wait.mean(), work.mean(), wait.mean(1).round(), work.mean(1).round()))
ew = f[:, :, 1] - f[:, :, 0]; sw = f[:, :, 2] - f[:, :, 1]; fw = f[:, :, 3] - f[:, :, 2]
print("fill: wait-for-empty %.0f, wait-for-stage %.0f, fill work %.0f clk per copy" % (ew.mean(), sw.mean(), fw.mean()))
per_copy = (g[:, 1:, 2] - g[:, :-1, 2]).mean()
print("steady-state clk per copy per gather warp: %.0f" % per_copy)
# lateness of the fill relative to the first gather warp that asks for it
ask = g[:, :, 0].min(0); full = f[:, :, 3].max(0)
print("fill completes %.0f clk after the first gather warp asks (mean over copies; negative = ready in time)" % (full - ask)[2:].mean())
