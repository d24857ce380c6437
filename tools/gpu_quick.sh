#!/usr/bin/env bash
# quick bench line: value ms/step k1 k0
python bench.py --steps ${1:-30} --warmup 3 --no-cpu-baseline --e2e-steps ${2:-0} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e  ms/step %.3f  k1 %.3f  k0 %.3f  frac %.3f  e2e %s' % (d['value'], d['ms_per_step'], r['k1_ms'], r['k0_ms'], r['frac'], d.get('e2e',{}).get('ms_per_step')))"
