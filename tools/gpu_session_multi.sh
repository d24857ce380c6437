#!/bin/bash
# tools/gpu_session_multi.sh <tag> <n_gpus> — the sharded / NCCL paths on N GPUs of one box.
TAG=${1:-r02m}; N=${2:-2}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1; nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
# small genome first (4 contigs = 40 windows): exercises sharding + gather + verification quickly
timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 3 --contigs 4 --e2e-windows 2 > $O/bench_c4_small_n$N.json 2> $O/bench_c4_small_n$N.err
echo "c4 small rc=$?"; tail -c 1500 $O/bench_c4_small_n$N.err
NCCL_MAX_NCHANNELS=4 timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 3 --contigs 4 --e2e-windows 0 --no-parity > $O/bench_c4_small_4ch_n$N.json 2> $O/bench_c4_small_4ch_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 3 --contigs 4 --e2e-windows 0 --no-parity --reserve-ctas 48 > $O/bench_c4_small_res_n$N.json 2> $O/bench_c4_small_res_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_c4_n$N.json 2> $O/bench_c4_n$N.err
echo "c4 rc=$?"; tail -c 800 $O/bench_c4_n$N.err
timeout 600 $TR bench.py --gpus $N --config c3 --steps 20 --warmup 3 > $O/bench_c3_n$N.json 2> $O/bench_c3_n$N.err
echo "c3 rc=$?"
timeout 600 $TR bench.py --gpus $N --config c5 --steps 2 --warmup 3 > $O/bench_c5_n$N.json 2> $O/bench_c5_n$N.err
echo "c5 rc=$?"; tail -c 400 $O/bench_c5_n$N.err
timeout 300 $TR bench.py --gpus $N --impl reference --steps 2 --warmup 1 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err
python - <<PY
import json
for f in ("bench_c4_small_n$N","bench_c4_small_4ch_n$N","bench_c4_small_res_n$N","bench_c4_n$N","bench_c3_n$N","bench_c5_n$N","bench_ref_n$N"):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1])
        print(f, "value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]), "e2e", d.get("e2e",{}).get("ms_per_step"), "parity", d.get("parity",{}).get("all_ranks_identical"), "nogather_ms", (d.get("config",{}).get("gather") or {}).get("ms_per_step_without_gather"), "ok", (d.get("config",{}).get("gather") or {}).get("verified_checksums"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
ls -la $O
