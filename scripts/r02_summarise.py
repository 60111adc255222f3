"""Turn the raw outputs of scripts/r02_measure.sh (gpurun_out/) into the summaries kept under profiles/:
   ncu --set full report -> profiles/r02_ncu_solve_kernels.txt + profiles/solve_traffic.json, launch list CSV ->
   profiles/r02_launch_list_bench64.txt, and copies of the JSON results.  Needs `ncu` (to read the .ncu-rep) but no GPU."""
import collections, csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

WANT = """dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed gpu__time_duration.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed launch__block_size launch__grid_size
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem launch__registers_per_thread launch__shared_mem_per_block_dynamic
lts__t_sector_hit_rate.pct lts__t_sectors.sum lts__throughput.avg.pct_of_peak_sustained_elapsed sm__cycles_elapsed.avg.per_second
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__warps_active.avg.pct_of_peak_sustained_active
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__warps_eligible.avg.per_cycle_active""".split()

raw = subprocess.run(["ncu", "-i", os.path.join(G, "r02_final_solve.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
out = ["# ncu --set full --clock-control none --import-source on, round 2, final kernels: python scripts/prof_solve.py 16 4",
       "# (16 images x 100 copies, 128^2->512^2); one launch of each solve kernel; cold-cache, serialised (compare shares, not absolutes)", ""]
traffic = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    short = "k_forward_residual" if "forward" in name else "k_gradient_update"
    out.append("## " + short + "   " + name[:110])
    for w in WANT:
        if w in ix:
            out.append(f"{w:90s} {r[ix[w]]:>18s} {units[ix[w]]}")
    out.append("")
    mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[ix["dram__bytes_read.sum"]]]
    traffic[short] = {"dram_bytes_per_image": (float(r[ix["dram__bytes_read.sum"]]) + float(r[ix["dram__bytes_write.sum"]])) * mult / 16}
open(os.path.join(P, "r02_ncu_solve_kernels.txt"), "w").write("\n".join(out))
json.dump({"source": "profiles/r02_ncu_solve_kernels.txt (ncu --set full, python scripts/prof_solve.py 16 4; dram__bytes_read.sum + "
                     "dram__bytes_write.sum of one launch / 16 images)", **traffic}, open(os.path.join(P, "solve_traffic.json"), "w"), indent=1)

rows = [r for r in csv.reader(open(os.path.join(G, "r02_launches_bench64.csv"))) if len(r) > 5]
hdr, data = None, []
for r in rows:
    if r[0] == "ID":
        hdr = r
    elif hdr and r[0].isdigit():
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in data:
    n = r[ix["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").replace("asr::", "")
    v, u = float(r[ix["Metric Value"]]), r[ix["Metric Unit"]]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
b = json.load(open(os.path.join(G, "r02_bench64_plain.json"))); r = b["roofline"]
lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_|asr -c 2000  python bench.py --images 64 --steps 1 --warmup 3 --no-cpu-baseline --no-l2-probe",
         "# first 2000 launches of libasr kernels (warm-up steps included); per-launch times are cold-cache and serialised: compare shares",
         "# kernel                       launches   total ms   share   avg us"]
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{n:30s} {c:8d} {t / 1e3:10.2f} {100 * t / tot:6.1f}% {t / c:9.1f}")
lines += ["", f"# same command without ncu (CUDA events on the launching stream): value {b['value']:.2f} images/s, K1 {r['forward_ms_total']:.1f} ms, "
              f"K2 {r['update_ms_total']:.1f} ms of the {b['ms_per_step']:.1f} ms timed step: K2 share {100 * r['update_ms_total'] / b['ms_per_step']:.1f}%, "
              f"K1 {100 * r['forward_ms_total'] / b['ms_per_step']:.1f}%"]
open(os.path.join(P, "r02_launch_list_bench64.txt"), "w").write("\n".join(lines) + "\n")
for f in ("r02_bench_1gpu.json", "r02_configs_1_3.json", "r02_configs_4_5.json", "r02_aux_kernels.json"):
    shutil.copy(os.path.join(G, f), os.path.join(P, f))
print("\n".join(lines[-12:]))
d = json.load(open(os.path.join(G, "r02_bench_1gpu.json"))); r = d["roofline"]
print("bench", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), r["us_per_image_iteration"], "frac", round(r["frac"], 4), "pair", round(r["pair_frac"], 4),
      "fp32", round(r["fp32_pipe_frac"], 4), "l2", round(r["l2"]["peak"]), round(r["l2"]["resident_run"]["value"], 1), d["cpu_baseline"]["value"])
print(json.load(open(os.path.join(G, "r02_configs_1_3.json"))))
