#!/usr/bin/env python
"""bench.py -- SR-solved images/s for BASELINE.json configs[1] (SR_single_class.py batch shape).

A step = one pass of the hot path over one batch: `--images` images per GPU (default 500), each with
100 augmented 128x128 copies solved to 512x512 by 300 Adam+AMSGrad iterations (test_SR.py / SR_single_class.py
hyper-parameters, shared-optimizer step offsets 300*j), then thresholded (th_factor 0.65, class 8).
Images shard across ranks with no collective in the solve; the masks are gathered to rank 0 (NCCL).

  value   whole-job images/s with the copies already resident in HBM (CUDA events, max over ranks)
  e2e     the same through the reference-facing Python API with HOST buffers: pinned H2D of the copies,
          solve, threshold, D2H of the masks, all inside the timed region
  roofline  dominant solve kernel, CUDA events on its launching stream, algorithmic bytes of SURVEY 8(d)
  cpu_baseline  the CPU oracle (a port of the reference's TensorFlow arithmetic) on this box's host cores

`--impl reference` times that CPU port alone (the reference itself is TensorFlow 2.7 code that cannot be
installed here: DESIGN.md, profiles/r02_tf_install_attempt.log), on a bounded sample of the same workload: the
first step is always one FULL 300-iteration solve of one image; later steps shorten the iteration count only if
the requested --steps would not fit the time budget.

`--scaling strong` runs BASELINE configs[1] as written: `--images` images IN TOTAL (default 500) sharded over the
ranks in contiguous blocks (sharding.shard_bounds); the default is weak scaling (`--images` per GPU).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "SR-solved images/sec (100 copies, 128^2->512^2)"
UNIT = "images/s"
ITERS, NUM_AUG, LR_HW, HR_HW = 300, 100, (128, 128), (512, 512)
# per image-iteration, read the LR-sized stack once + read/write x + read/write m, v, vhat
BYTES_PER_IMAGE_ITER = 4 * NUM_AUG * LR_HW[0] * LR_HW[1] + 8 * 4 * HR_HW[0] * HR_HW[1]
# DESIGN.md "Roofline": un-fusable fp32 operations the literal operator sequence needs per image-iteration --
# 22 per (HR pixel x copy) gradient gather (4 coordinate adds, 4 floor, 4 weights, 9 lerp, 1 accumulate),
# 21 per forward rotate gather (9 per LR cell x copy), 29 per LR cell x copy for translate+resize+residual,
# 21 per LR cell x copy to expand the residual into the translate-gradient image
FP32_OPS_PER_IMAGE_ITER = NUM_AUG * (22 * HR_HW[0] * HR_HW[1] + (9 * 21 + 29 + 21) * LR_HW[0] * LR_HW[1])
FP32_LANES_PER_SM = 128
N_SM = 148


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=500, help="images per GPU per step (weak scaling)")
    ap.add_argument("--images-in-flight", type=int, default=0, help="images per kernel-launch group (0 = all)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --images per GPU; strong: --images in total")
    ap.add_argument("--cpu-seconds", type=float, default=45.0, help="budget of the cpu_baseline sample (one full solve when it fits)")
    ap.add_argument("--ref-seconds", type=float, default=240.0, help="--impl reference: wall-clock budget of the whole run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-probe", action="store_true", help="skip the L2 bandwidth microbenchmark and the L2-resident run")
    ap.add_argument("--l2-images", type=int, default=4, help="images in flight of the L2-resident run (about 20 MB of working set each)")
    return ap.parse_args()


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


# ---------------------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=10)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU port of the reference (oracle) on a bounded sample
# ---------------------------------------------------------------------------------------------------------------
def oracle_sample(copies_np, ang, sh, iters):
    from oracle import oracle as O
    O.use_all_cores()
    P = O.SolveParams(num_iter=int(iters))
    t = time.perf_counter()
    O.augmented_superresolution(copies_np, ang, sh, P, output_size=HR_HW)
    dt = time.perf_counter() - t
    return dt, O.num_threads()


def cpu_baseline(copies_np, ang, sh, budget_s):
    dt3, cores = oracle_sample(copies_np, ang, sh, 3)
    per_iter = dt3 / 3
    n = ITERS if per_iter * ITERS <= budget_s else int(max(5, budget_s / per_iter))
    dt, cores = oracle_sample(copies_np, ang, sh, n)
    per_iter = dt / n
    how = "one full solve" if n == ITERS else "extrapolated linearly in iterations"
    return {"value": 1.0 / (per_iter * ITERS), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 image x {NUM_AUG} copies x {n} of {ITERS} iterations in {dt:.1f} s ({how})"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
    copies, ang, sh = make_augmented_copies(1, NUM_AUG, LR_HW, HR_HW, 0.15, 80, seed=1234, value=1.0)
    c = copies[0].numpy()
    n_steps = args.warmup + args.steps
    # step 0 (a warm-up unless --warmup 0) is ONE FULL solve whenever half the budget can pay for it; it also calibrates what
    # the remaining steps can afford
    t3, cores = oracle_sample(c, ang[0], sh[0], 3)
    full = (t3 / 3) * ITERS <= 0.5 * args.ref_seconds
    t_full, cores = oracle_sample(c, ang[0], sh[0], ITERS) if full else ((t3 / 3) * ITERS, cores)
    left = max(1, n_steps - 1)
    iters = ITERS if t_full * left <= max(0.0, args.ref_seconds - t_full) else int(max(6, ITERS * (args.ref_seconds - t_full) / (t_full * left)))
    iters = min(ITERS, max(6, iters))
    times = [t_full * iters / ITERS] if args.warmup == 0 else []      # per-step times, normalised to `iters` iterations
    for s_ in range(1, n_steps):
        dt, cores = oracle_sample(c, ang[0], sh[0], iters)
        if s_ >= args.warmup:
            times.append(dt)
    per_iter = sum(times) / (len(times) * iters)
    value = 1.0 / (per_iter * ITERS)
    if full:
        head = (f"step 0 = 1 image x {NUM_AUG} copies x {ITERS} of {ITERS} iterations (one full solve, {t_full:.1f} s = "
                f"{1.0 / t_full:.5f} images/s); ")
    else:
        head = (f"a full solve would take ~{t_full:.0f} s on this host (3-iteration probe), more than half of --ref-seconds, "
                f"so none was run; ")
    sample = head + f"timed steps = 1 image x {iters} of {ITERS} iterations each" + \
        ("" if iters == ITERS else ", images/s extrapolated linearly in iterations")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "configs[1] SR_single_class batch shape: 100 copies, 128^2->512^2, 300 Adam+AMSGrad iterations",
                       "note": "CPU port (oracle/asr_oracle.c) of the reference's TensorFlow op sequence on all host cores; the reference "
                               "itself needs tensorflow==2.7.0 + tensorflow-addons==0.15.0, which cannot be installed here "
                               "(profiles/r02_tf_install_attempt.log)", "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "full_solve_s": t_full if full else None},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
def emit(line):
    """The one JSON line goes to the real stdout; everything else (NCCL's version banner included) was sent to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import ctypes as C
    import torch
    import torch.distributed as dist
    from deeplabv3plus_augmented_superresolution_b200 import _lib as A, sharding
    from deeplabv3plus_augmented_superresolution_b200.synthetic import make_augmented_copies
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.optimizer import Optimizer
    from deeplabv3plus_augmented_superresolution_b200.superresolution_scripts.superresolution import Superresolution

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libasr has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = A.lib()

    if args.scaling == "strong":
        n_total = args.images                                   # configs[1] as written: 500 images over all GPUs
        lo, hi = sharding.shard_bounds(n_total, world, rank)
    else:
        n_total = args.images * world                            # the same per-GPU batch on every rank
        lo, hi = rank * args.images, (rank + 1) * args.images
    B = hi - lo
    if B <= 0:
        raise SystemExit(f"rank {rank} has no images: --images {args.images} over {world} ranks")
    # synthetic stand-in for the hdf5 augmented-copies files: rank r holds images [lo, hi) of the run
    copies, ang, sh = make_augmented_copies(B, NUM_AUG, LR_HW, HR_HW, 0.15, 80, seed=1234 + 7 * rank, value=1.0, device=dev)

    def make_solver():
        opt = Optimizer(optimizer="adam", learning_rate=1e-3, amsgrad=True, lr_scheduler=True, decay_steps=60, decay_rate=0.3)
        return Superresolution(lambda_df=1.0, lambda_tv=0.3, lambda_L2=0.7, lambda_L1=0.0, num_iter=ITERS, num_aug=NUM_AUG,
                               optimizer=opt, feature_size=LR_HW, output_size=HR_HW)

    ws_th = torch.empty(2 * B, dtype=torch.float32, device=dev)
    masks = torch.empty((B, HR_HW[0], HR_HW[1]), dtype=torch.int32, device=dev)

    def hot_path(dev_copies, in_flight=None, n=None):
        n = B if n is None else n
        sr = make_solver()
        sr.optimizer.iterations = ITERS * lo                     # one optimizer shared by the whole (sharded) run
        fl = args.images_in_flight if in_flight is None else in_flight
        plist = [sr._solve_params(sr.optimizer.iterations + j * ITERS, images_in_flight=fl) for j in range(n)]
        x = sr.augmented_superresolution_batched(dev_copies[:n], ang[:n], sh[:n], params_list=plist)
        A.check(L.asr_threshold(x.data_ptr(), n, HR_HW[0] * HR_HW[1], 8, 0.65, None, masks.data_ptr(), ws_th.data_ptr(),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        if n != B:
            return masks[:n]
        return sharding.gather_masks(masks.to(torch.uint8), n_total) if world > 1 else masks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        hot_path(copies)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    L.asr_profile_enable(1)
    launches0 = L.asr_kernel_launches()
    ms = timed(lambda: hot_path(copies), args.steps)
    launches = L.asr_kernel_launches() - launches0
    L.asr_profile_enable(0)
    kms = (C.c_double * 2)(); kcnt = (C.c_longlong * 2)()
    A.check(L.asr_profile_read(kms, kcnt))
    clocks = sampler.stop() if sampler else None
    value = n_total * args.steps / (ms / 1e3)

    # ---- end-to-end arm: host buffers in, host masks out -----------------------------------------------
    host_copies = torch.empty(copies.shape, dtype=torch.float32, pin_memory=True)
    host_copies.copy_(copies)
    host_masks = torch.empty(masks.shape, dtype=torch.int32, pin_memory=True)
    stage = torch.empty_like(copies)

    def e2e_step():
        stage.copy_(host_copies, non_blocking=True)
        hot_path(stage)
        host_masks.copy_(masks, non_blocking=True)

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    bytes_io = torch.tensor([host_copies.numel() * 4, host_masks.numel() * 4, launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(bytes_io)
    e2e = {"value": n_total * args.steps / (ms_e2e / 1e3), "unit": UNIT,
           "h2d_bytes_per_step": int(bytes_io[0].item()), "d2h_bytes_per_step": int(bytes_io[1].item()),
           "api": "Superresolution.augmented_superresolution_batched (the batched extension of the reference-named class; the reference's own "
                  "per-image signature is measured in profiles/ as config 1) + asr_threshold, pinned host buffers in and out"}

    # ---- L2 regime (rank 0, one GPU): measured L2 read bandwidth, and the solve with so few images in flight that
    #      their whole working set (copies + residuals + x + slots + tap tables, ~20 MB per image) stays in the 126 MB L2
    l2 = None
    if rank == 0 and world == 1 and not args.no_l2_probe:
        nb = 64 << 20
        buf = torch.empty(nb, dtype=torch.uint8, device=dev).fill_(1)
        sink = torch.zeros(1, dtype=torch.int32, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        A.check(L.asr_l2_read_probe(buf.data_ptr(), nb, 2, sink.data_ptr(), st))            # warm
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            passes = 40
            e0.record(); A.check(L.asr_l2_read_probe(buf.data_ptr(), nb, passes, sink.data_ptr(), st)); e1.record()
            torch.cuda.synchronize()
            best = max(best, nb * passes / (e0.elapsed_time(e1) / 1e3) / 1e9)
        del buf
        n_l2 = min(B, 8 * args.l2_images)
        hot_path(copies, in_flight=args.l2_images, n=n_l2)
        ms_l2 = timed(lambda: hot_path(copies, in_flight=args.l2_images, n=n_l2), 1)
        ips = n_l2 / (ms_l2 / 1e3)
        ach = ips * ITERS * BYTES_PER_IMAGE_ITER / 1e9
        l2 = {"peak": best, "unit": "GB/s", "peak_source": "asr_l2_read_probe: 64 MB buffer read 40x with ld.global.cg, best of 3, CUDA events",
              "resident_run": {"images_in_flight": args.l2_images, "images": n_l2, "value": ips, "unit": UNIT, "achieved": ach, "frac": ach / best,
                               "note": "algorithmic bytes of SURVEY 8(d) / time with the working set of the launch group L2-resident; "
                                       "ncu lts__throughput of the two kernels: profiles/r02_ncu_solve_kernels.txt"}}

    if rank == 0:
        peaks, peak_kind = load_peaks()
        dom = 1 if kms[1] >= kms[0] else 0
        names = ["k_forward_residual", "k_gradient_update"]
        avg_launch_s = (kms[dom] / max(1, kcnt[dom])) / 1e3
        # launches of a short tail group carry fewer images: use the true mean images per launch (rank 0's share)
        images_per_launch = B * ITERS * args.steps / max(1, kcnt[dom])
        achieved = images_per_launch * BYTES_PER_IMAGE_ITER / avg_launch_s / 1e9
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "solve_traffic.json")))
            traffic = prof[names[dom]]["dram_bytes_per_image"] * images_per_launch
        except Exception:
            pass
        iter_s = (kms[0] + kms[1]) / 1e3 / max(1, kcnt[1])
        pair_gbs = images_per_launch * BYTES_PER_IMAGE_ITER / iter_s / 1e9
        # the bound that actually binds: fp32 lane-operations per second against 148 SMs x 128 lanes x the SM clock under load
        clk_hz = 1e6 * float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0)
        fp_peak = N_SM * FP32_LANES_PER_SM * clk_hz
        fp_ach = images_per_launch * FP32_OPS_PER_IMAGE_ITER / iter_s
        roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "frac_is": f"{names[dom]} alone (the dominant kernel); the whole iteration is pair_frac",
                    "traffic": traffic, "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
                    "share_of_step": kms[dom] / ms, "avg_launch_ms": avg_launch_s * 1e3,
                    "forward_ms_total": kms[0], "update_ms_total": kms[1],
                    "us_per_image_iteration": {"k_forward_residual": kms[0] * 1e3 / max(1, kcnt[0]) / images_per_launch,
                                               "k_gradient_update": kms[1] * 1e3 / max(1, kcnt[1]) / images_per_launch},
                    "pair_achieved": pair_gbs, "pair_frac": pair_gbs / peaks["hbm_gbs"],
                    "fp32_pipe_frac": fp_ach / fp_peak,
                    "fp32_pipe": {"achieved": fp_ach / 1e12, "peak": fp_peak / 1e12, "unit": "T fp32 lane-op/s", "frac": fp_ach / fp_peak,
                                  "ops_per_image_iteration": FP32_OPS_PER_IMAGE_ITER,
                                  "note": "un-fused IEEE ops of the literal operator sequence (no FMA by contract); K1+K2 launch pair"},
                    "l2": l2,
                    "note": "the solve is FP32-issue bound, not bandwidth bound (pair_frac is small by construction): DESIGN.md 'Roofline', profiles/"}
        workload = (f"configs[1] SR_single_class batch: {args.images} images x 100 copies " +
                    ("per GPU" if args.scaling == "weak" else f"in total, sharded over {world} GPU(s)") +
                    ", 128^2->512^2, 300 Adam+AMSGrad iterations (test_SR.py hyper-parameters, shared-optimizer step offsets), threshold 0.65")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": workload, "images_total": n_total, "images_rank0": B, "num_aug": NUM_AUG, "iterations": ITERS,
                           "images_in_flight": args.images_in_flight,
                           "parallelism": f"images sharded over {world} GPU(s), no collective in the solve, NCCL gather of masks",
                           "l2": f"inputs are {copies.numel() * 4 / 1e9:.2f} GB on rank 0 (> 126 MB L2), no flush needed",
                           "parity": "bit-identical to the CPU oracle (tests/test_parity_gpu.py); the oracle is a port, not TensorFlow (DESIGN.md 'Oracle')"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(bytes_io[2].item()), "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(copies[0].cpu().numpy(), ang[0], sh[0], args.cpu_seconds)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
