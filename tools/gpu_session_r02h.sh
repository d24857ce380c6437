#!/bin/bash
# tools/gpu_session_r02h.sh <tag> — 1 GPU: parity tests, then A/B of the tile dispenser and the two-handle e2e caller
TAG=${1:-r02h}; O=gpurun_out/$TAG; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) > $O/pytest.log 2>&1; tail -2 $O/pytest.log
B="--config c3 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e-text --no-parity"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $B > $O/c3_$name.json 2> $O/c3_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/c3_$name.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("c3 $name", "ms/step %.3f k1 %.3f k0 %.3f frac %.3f" % (d["ms_per_step"], r["k1_ms"], r["k0_ms"], r["frac"]), "e2e %.2f" % d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"])
except Exception as ex: print("c3 $name FAILED", ex, open("$O/c3_$name.err").read()[-500:])
PY
}
run dyn X=1
run static BRC_K1_STATIC_TILES=1
C="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e-text --no-parity --contigs 6"
c4() { name=$1; shift; timeout 400 python bench.py $C "$@" > $O/c4_$name.json 2> $O/c4_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/c4_$name.json").read().strip().splitlines()[-1])
    print("c4 $name", "value %.4g ms/step %.2f" % (d["value"], d["ms_per_step"]), "e2e ms %.2f value %.4g" % (d["e2e"]["ms_per_step"], d["e2e"]["value"]), "handles", d["e2e"].get("handles_in_flight"), "h2d", d["e2e"]["h2d_bytes_per_step"], "d2h", d["e2e"]["d2h_bytes_per_step"])
except Exception as ex: print("c4 $name FAILED", ex, open("$O/c4_$name.err").read()[-500:])
PY
}
c4 h1 --e2e-handles 1
c4 h2 --e2e-handles 2
c4 h3 --e2e-handles 3
