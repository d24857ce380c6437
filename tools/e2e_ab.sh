#!/usr/bin/env bash
# A/B of the push path on one box: early (speculative) H2D vs H2D issued inside brc_compute
for mode in early late early late; do
  if [ $mode = late ]; then export BRC_NO_EARLY_H2D=1; else unset BRC_NO_EARLY_H2D; fi
  BRC_PIPE_TIMING=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 4 2> /tmp/e2e_$mode.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$mode', 'e2e ms', round(d['e2e']['ms_per_step'],2))"
  grep "brc pipe\] chunks" /tmp/e2e_$mode.err | tail -1
done
