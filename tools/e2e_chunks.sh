#!/usr/bin/env bash
for c in 2 4 8 16 32 64; do
  BRC_PIPE_CHUNKS=$c BRC_PIPE_TIMING=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 4 2> /tmp/e2e_c.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunks $c', 'e2e ms', round(d['e2e']['ms_per_step'],2))"
  grep "brc pipe\] chunks" /tmp/e2e_c.err | tail -1
done
