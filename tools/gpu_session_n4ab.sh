#!/bin/bash
# tools/gpu_session_n4ab.sh <tag> — 4 GPUs: sharded C4 with the tile dispenser vs the fixed stride, then 2 of the 4 GPUs
TAG=${1:-r02n4}; O=gpurun_out/$TAG; mkdir -p $O
tr() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
F="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e-text --no-parity --e2e-windows 0"
tr 4 $F > $O/c4_n4_dyn.json 2> $O/c4_n4_dyn.err; echo "n4 dyn rc=$?"
BRC_K1_STATIC_TILES=1 tr 4 $F > $O/c4_n4_static.json 2> $O/c4_n4_static.err; echo "n4 static rc=$?"
tr 2 $F > $O/c4_n2_dyn.json 2> $O/c4_n2_dyn.err; echo "n2 dyn rc=$?"
tr 4 --config c5 --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $O/c5_n4.json 2> $O/c5_n4.err; echo "c5 n4 rc=$?"
python - <<PY
import json
for f in ("c4_n4_dyn","c4_n4_static","c4_n2_dyn","c5_n4"):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1]); g=d.get("config",{}).get("gather") or {}
        print(f, "value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]), "nogather_ms", g.get("ms_per_step_without_gather"), "ok", g.get("verified_checksums"), "ingress", g.get("rank0_ingress_probe_GBps"))
    except Exception as ex:
        print(f, "FAILED", ex, open("$O/"+f+".err").read()[-400:])
PY
