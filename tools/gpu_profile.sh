#!/usr/bin/env bash
# Run on the GPU box (under gpurun): launch list + full ncu capture of the two kernels.
set -uo pipefail
TAG="${1:-r01}"
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pileup_kernel -s 3 -c 1 -f -o gpurun_out/prof_k1_${TAG} $BENCH >> gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:read_precompute -s 3 -c 1 -f -o gpurun_out/prof_k0_${TAG} $BENCH >> gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ls -la gpurun_out
