"""Config 3 (BASELINE.json): augmentation warp + OPM extraction (+ normalise / threshold / back-projection)
on synthetic inputs of the reference shapes; CUDA-event timing, algorithmic bytes, fraction of the
measured HBM peak.  Writes one JSON object (also used to fill profiles/)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deeplabv3plus_augmented_superresolution_b200 import _lib as A
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts import augmentation_utils as AU, superres_utils as SU
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.optimizer import Optimizer
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()                       # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


out = {}
N = 100
g = torch.Generator(device="cuda").manual_seed(0)
img = torch.rand((512, 512, 3), device="cuda", generator=g)
np.random.seed(1234)
ang, sh = AU._draw(N, 0.15, 80)
for interp in ("bilinear", "nearest"):
    ms = timeit(lambda: AU.warp_copies(img, ang, sh, interp))
    b = N * 512 * 512 * 3 * 4 + 512 * 512 * 3 * 4
    out[f"warp_affine_{interp}"] = {"ms": ms, "algorithmic_bytes": b, "GBps": b / ms / 1e6, "frac_hbm": b / ms / 1e6 / peak}
logits = torch.randn((N, 128, 128, 21), device="cuda", generator=g)
for mode in ("argmax", "slice", "slice_max"):
    ms = timeit(lambda: AU.extract_opm(logits, 8, mode))
    b = logits.numel() * 4 * (2 if mode == "slice" else 1) + N * 128 * 128 * 4 * (2 if mode == "slice_max" else 1)
    out[f"opm_{mode}"] = {"ms": ms, "algorithmic_bytes": b, "GBps": b / ms / 1e6, "frac_hbm": b / ms / 1e6 / peak}
# the same OPM extraction for 8 images' worth of predictions in one launch (the batch_runner shape): at 100 copies the kernel is
# launch-latency sized (~40 us), the bandwidth fraction only means something with more work per launch
logits8 = torch.randn((8 * N, 128, 128, 21), device="cuda", generator=g)
for mode in ("argmax", "slice", "slice_max"):
    ms = timeit(lambda: AU.extract_opm(logits8, 8, mode))
    b = logits8.numel() * 4 * (2 if mode == "slice" else 1) + 8 * N * 128 * 128 * 4 * (2 if mode == "slice_max" else 1)
    out[f"opm_{mode}_800copies"] = {"ms": ms, "algorithmic_bytes": b, "GBps": b / ms / 1e6, "frac_hbm": b / ms / 1e6 / peak}
del logits8
copies, a2, s2 = make_augmented_copies(8, N, device="cuda")
stack = copies[0, :, :, :, None].contiguous()
ms = timeit(lambda: SU._normalize_stack_device(stack))
b = stack.numel() * 4 * 3
out["minmax_normalize"] = {"ms": ms, "algorithmic_bytes": b, "GBps": b / ms / 1e6, "frac_hbm": b / ms / 1e6 / peak}
sr = Superresolution(1, 0.3, 0.7, 0, optimizer=Optimizer(), feature_size=(128, 128))
for mode in ("max", "mean"):
    ms = timeit(lambda: sr.backproject_batched(copies, a2, s2, mode), reps=5)
    b = copies.numel() * 4 + 8 * 512 * 512 * 4
    out[f"backproject_{mode}_8img"] = {"ms": ms, "algorithmic_bytes": b, "GBps": b / ms / 1e6, "frac_hbm": b / ms / 1e6 / peak, "images_per_s": 8 / ms * 1e3}
x = torch.rand((64, 512, 512), device="cuda", generator=g)
o = torch.empty((64, 512, 512), dtype=torch.int32, device="cuda"); ws = torch.empty(128, device="cuda")
import ctypes as C
L = A.lib()
ms = timeit(lambda: A.check(L.asr_threshold(x.data_ptr(), 64, 512 * 512, 8, 0.65, None, o.data_ptr(), ws.data_ptr(), None)))
b = x.numel() * 4 * 3
out["threshold_64img"] = {"ms": ms, "algorithmic_bytes": b, "GBps": b / ms / 1e6, "frac_hbm": b / ms / 1e6 / peak}
print(json.dumps({"hbm_peak_GBps": peak, "kernels": out}, indent=1))
