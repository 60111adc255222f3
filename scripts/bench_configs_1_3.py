"""BASELINE.json configs 1 and 3 as measurements through the reference-facing Python API (host arrays in, host arrays out).
   config 1: test_SR.py path: 1 synthetic image, 100 copies {0,8}, 128^2 -> 512^2, 300 Adam+AMSGrad iterations, threshold 0.2, class 8
   config 3: max-OPM pipeline on random logits: warp 100 RGB copies, slice_max OPM extraction, two solves (class and max maps),
             threshold class >= max
Prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.optimizer import Optimizer
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution
from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts import augmentation_utils as AU, superres_utils as SU



def solver():
    opt = Optimizer(optimizer="adam", learning_rate=1e-3, amsgrad=True, lr_scheduler=True, decay_steps=60, decay_rate=0.3)
    return Superresolution(lambda_df=1.0, lambda_tv=0.3, lambda_L2=0.7, lambda_L1=0.0, num_iter=300, num_aug=100, optimizer=opt,
                           feature_size=(128, 128), output_size=(512, 512))


def wall(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return float(np.median(ts))


out = {}
# ---- config 1 ----------------------------------------------------------------------------------------------
copies, ang, sh = make_augmented_copies(1, 100, (128, 128), (512, 512), 0.15, 80, seed=1234, value=8.0)
clist = [c[..., None] for c in copies[0].numpy()]          # the list of [h,w,1] arrays test_SR.py builds
def config1():
    x, _ = solver().augmented_superresolution(clist, ang[0], sh[0])
    return SU.threshold_image(x, 8, th_factor=0.2)
s1 = wall(config1)
# the CPU side of this config is bench.py's cpu_baseline (the oracle is test infrastructure: only tests/, smoke() and
# bench.py's cpu_baseline / --impl reference legs execute it); pass its images/s as argv[1] to get the ratio
cpu_ips = float(sys.argv[1]) if len(sys.argv) > 1 else None
out["config1_test_SR"] = {"gpu_seconds_per_image": s1, "gpu_images_per_s": 1 / s1}
if cpu_ips:
    out["config1_test_SR"].update({"cpu_oracle_images_per_s_from_bench_py": cpu_ips, "speedup": (1 / s1) / cpu_ips})
# ---- config 3 ----------------------------------------------------------------------------------------------
g = torch.Generator(device="cuda").manual_seed(0)
img = torch.rand((512, 512, 3), device="cuda", generator=g).cpu().numpy()
logits = torch.randn((100, 128, 128, 21), device="cuda", generator=g)
np.random.seed(1234)
def config3():
    warped, a, s = AU.create_augmented_copies(img, 100, 0.15, 80)          # [100,512,512,3] on the device (DeepLabV3+ would consume it)
    cls, mx = AU.extract_opm(logits, 8, "slice_max")
    sr = solver()
    xc = sr.augmented_superresolution_batched(torch.stack([cls[..., 0], mx[..., 0]]), np.stack([a, a]), np.stack([s, s]))
    return SU.threshold_image(xc[0].cpu().numpy()[..., None], 8, th_mask=xc[1].cpu().numpy()[..., None])
s3 = wall(config3)
out["config3_max_opm_pipeline"] = {"gpu_seconds_per_image": s3, "stages": "warp 100 RGB copies + slice_max OPM + two solves in one batch + class>=max threshold"}
print(json.dumps(out, indent=1))
