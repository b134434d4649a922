#!/bin/bash
# Multi-GPU evidence run on one box (gpurun --gpus 8): in-kernel NVLink exchange parity at 2/4/8 ranks, K2 strong scaling,
# config 2 / 3 weak scaling.  Set APS_MG_FULL=1 to add the 4-GPU K2 line and config 4.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_k2.py tests/test_multi_gpu.py -m gpu -q -k "multi_gpu or sharded" > gpurun_out/r2_pytest_8gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_8gpu.log
N=8; T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
$T bench.py --gpus $N --workload k2 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2_bench_k2_${N}gpu.err | grep "^{" > gpurun_out/r2_bench_k2_${N}gpu.json; echo "k2 N=$N rc=$?"; cut -c1-200 gpurun_out/r2_bench_k2_${N}gpu.json
$T bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-k2 2> gpurun_out/r2_bench_config2_${N}gpu.err | grep "^{" > gpurun_out/r2_bench_config2_${N}gpu.json; echo "config2 N=$N rc=$?"; cut -c1-200 gpurun_out/r2_bench_config2_${N}gpu.json
$T bench.py --gpus $N --workload config3 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2_bench_config3_${N}gpu.err | grep "^{" > gpurun_out/r2_bench_config3_${N}gpu.json; echo "config3 N=$N rc=$?"; cut -c1-200 gpurun_out/r2_bench_config3_${N}gpu.json
if [ "${APS_MG_FULL:-0}" = "1" ]; then
  $T bench.py --gpus $N --workload config4 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2_bench_config4_${N}gpu.err | grep "^{" > gpurun_out/r2_bench_config4_${N}gpu.json; echo "config4 N=$N rc=$?"
  N=4; T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534"
  $T bench.py --gpus $N --workload k2 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2_bench_k2_${N}gpu.err | grep "^{" > gpurun_out/r2_bench_k2_${N}gpu.json; echo "k2 N=$N rc=$?"
fi
