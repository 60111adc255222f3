"""BASELINE.json configs 4 and 5 as measurements (the headline bench covers config 2, tests cover parity):
   config 4: num_aug sweep 16/64/256/1024 copies per image (L2-resident vs HBM-spill LR stacks)
   config 5: hyper-parameter grid (TV lambda x learning rate x iterations) over images, many solves in flight,
             every point reading its image's LR stack in place (asr_solve_sweep)
Prints one JSON object."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplabv3plus_augmented_superresolution_b200 import _lib as A
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies

peak = 6534.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3


out = {"hbm_peak_GBps": peak, "config4_num_aug": [], "config5_sweep": None}
for n_aug, B in ((16, 128), (64, 64), (100, 64), (256, 32), (1024, 8)):
    copies, ang, sh = make_augmented_copies(B, n_aug, device="cuda")
    s = timed(lambda: A.solve_batched(copies, ang, sh, A.SolveParams(num_iter=300)))
    bytes_iter = 4 * n_aug * 128 * 128 + 8 * 4 * 512 * 512
    ips = B / s
    out["config4_num_aug"].append({"num_aug": n_aug, "images": B, "images_per_s": ips, "lr_stack_MB": 4 * n_aug * 128 * 128 / 1e6,
                                   "us_per_image_iteration": s / B / 300 * 1e6, "us_per_copy_iteration": s / B / 300 / n_aug * 1e6,
                                   "frac_bandwidth_roofline": ips * 300 * bytes_iter / 1e9 / peak})
    del copies
images, grid = 20, [dict(lambda_tv=tv, learning_rate=lr, num_iter=it) for tv in (0.1, 0.3, 0.5) for lr in (5e-4, 1e-3) for it in (200, 300)]
copies, ang, sh = make_augmented_copies(images, 100, device="cuda")
points = [(s, g) for s in range(images) for g in grid]
plist = [A.SolveParams(**g) for _, g in points]
idx = [s for s, _ in points]
s = timed(lambda: A.solve_sweep(copies, ang, sh, plist, idx))
iters_total = sum(p.num_iter for p in plist)
out["config5_sweep"] = {"images": images, "points_per_image": len(grid), "solves": len(points), "seconds": s, "solves_per_s": len(points) / s,
                        "image_iterations_per_s": iters_total / s, "equiv_images_per_s_at_300_iterations": iters_total / 300 / s}
print(json.dumps(out, indent=1))
