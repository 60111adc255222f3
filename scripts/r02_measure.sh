#!/bin/bash
# Round-2 measurement set (one B200): full bench, configs 1/3/4/5, aux kernels, ncu launch list and full capture of the solve kernels.
set -x
python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
CPU=$(python -c "import json; print(json.load(open('gpurun_out/r02_bench_1gpu.json'))['cpu_baseline']['value'])")
python scripts/bench_configs_1_3.py $CPU > gpurun_out/r02_configs_1_3.json 2> gpurun_out/r02_c13.err
python scripts/bench_configs.py > gpurun_out/r02_configs_4_5.json 2> gpurun_out/r02_c45.err
python scripts/bench_aux.py > gpurun_out/r02_aux_kernels.json 2> gpurun_out/r02_aux.err
# launch list of a bench command that already exited 0 without ncu (per-launch times are cold-cache and serialised: compare shares)
python bench.py --images 64 --steps 1 --warmup 3 --no-cpu-baseline --no-l2-probe > gpurun_out/r02_bench64_plain.json 2> gpurun_out/r02_b64.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|asr' -c 2000 --csv --log-file gpurun_out/r02_launches_bench64.csv \
    python bench.py --images 64 --steps 1 --warmup 3 --no-cpu-baseline --no-l2-probe > gpurun_out/r02_ncu_launch.log 2>&1
python scripts/prof_solve.py 16 4 > gpurun_out/r02_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_forward_residual|k_gradient_update' -s 4 -c 2 -f -o gpurun_out/r02_final_solve \
    python scripts/prof_solve.py 16 4 > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out/r02_final_solve.ncu-rep
