#!/bin/bash
# tools/gpu_session_n8.sh <tag> — the sharded C4 / C5 paths on 8 GPUs of one box, then C4 again on 4 and 2 of them (same box, same build)
TAG=${1:-r02m8}; O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1; nvidia-smi topo -m > $O/topo.txt 2>&1
tr() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
tr 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e-text > $O/bench_c4_n8.json 2> $O/bench_c4_n8.err; echo "c4 n8 rc=$?"; tail -c 600 $O/bench_c4_n8.err
tr 8 --config c5 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_c5_n8.json 2> $O/bench_c5_n8.err; echo "c5 n8 rc=$?"
tr 4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e-text --no-parity --e2e-windows 0 > $O/bench_c4_n4.json 2> $O/bench_c4_n4.err; echo "c4 n4 rc=$?"
tr 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e-text --no-parity --e2e-windows 0 > $O/bench_c4_n2.json 2> $O/bench_c4_n2.err; echo "c4 n2 rc=$?"
python - <<PY
import json
for f in ("bench_c4_n8","bench_c5_n8","bench_c4_n4","bench_c4_n2"):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1]); g=d.get("config",{}).get("gather") or {}
        print(f, "value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]), "e2e", d.get("e2e",{}).get("ms_per_step"), d.get("e2e",{}).get("value"), "parity", d.get("parity",{}).get("all_ranks_identical"),
              "nogather_ms", g.get("ms_per_step_without_gather"), "ok", g.get("verified_checksums"), "ingress", g.get("rank0_ingress_probe_GBps"), "numa", d["config"].get("numa"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
