#!/bin/bash
# Round-end evidence run on one B200 (gpurun): tests, bench lines, ncu launch list and full captures.  Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_final.log; tail -4 gpurun_out/r2_pytest_final.log
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_final_reference_arm.json 2> gpurun_out/r2_final_reference_arm.err; echo "ref rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_final_bench_1gpu.json
for w in config3 config4 k2; do python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/r2_final_bench_$w.json 2> gpurun_out/r2_final_bench_$w.err; echo "$w rc=$?"; done
# launch list of the bench command (per-launch times are cold-cache and serialised: only the kernel's SHARE of the step is meaningful)
APS_BENCH_NO_SAMPLER=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-k2 > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu list rc=$?"
# full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:k1_lean -c 1 -o gpurun_out/r2_k1_lean python tools/quick_bench.py --replicas 4144 --T 2 --reps 1 > gpurun_out/r2_ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k2_pass -s 12 -c 1 -o gpurun_out/r2_k2_local_final python tools/k2_bench.py --logL 26 --case "local sigma=5 dt=0.005" --passes 20 > gpurun_out/r2_ncu_k2l.log 2>&1; echo "ncu k2 local rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k2_pass -s 12 -c 1 -o gpurun_out/r2_k2_global_final python tools/k2_bench.py --logL 26 --case "global dt=0.02" --passes 20 > gpurun_out/r2_ncu_k2g.log 2>&1; echo "ncu k2 global rc=$?"
python tools/k2_dt_bias.py --replicas 768 --out gpurun_out/r2_k2_dt_bias.json > gpurun_out/r2_k2_dt_bias.md 2> gpurun_out/r2_k2_dt_bias.err; cat gpurun_out/r2_k2_dt_bias.md
ls -la gpurun_out/*.ncu-rep
