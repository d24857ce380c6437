#!/bin/bash
# tools/gpu_session_e2e.sh <tag> — PCIe pipeline experiments on the C3 window (e2e leg only) + the 4-GPU sharded run is separate
TAG=${1:-r02e}; O=gpurun_out/$TAG; mkdir -p $O
B="--config c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e-text --no-parity"
run() { name=$1; shift; env "$@" BRC_PIPE_TIMING=1 timeout 300 python bench.py $B > $O/e2e_$name.json 2> $O/e2e_$name.err; echo "$name: $(python -c "import json;d=json.loads(open('$O/e2e_$name.json').read().strip().splitlines()[-1]);print('e2e %.2f ms' % d['e2e']['ms_per_step'])") | $(grep 'device clocks' $O/e2e_$name.err | tail -1)"; }
run base X=1
run d2h1d BRC_D2H_1D=1
run chunks2 BRC_PIPE_CHUNKS=2
run chunks4 BRC_PIPE_CHUNKS=4
run chunks16 BRC_PIPE_CHUNKS=16
run chunks32 BRC_PIPE_CHUNKS=32
run early BRC_EARLY_H2D=1
run two BRC_H2D_TWO_STREAMS=1
