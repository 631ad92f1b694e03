#!/bin/bash
# A/B of bench.py under environment switches: prints ms_per_step (two-stream timed region), e2e, serial, recurrence
# usage: bench_ab.sh "VAR=val ..." ["VAR=val ..."] ...
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --no-cpu --no-sampler --no-extras --steps 20 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$cfg', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'serial', round(d['schedule']['serial_ms_per_step'],3), 'cats', {k: round(v,3) for k,v in d['per_category_ms'].items()})"
done
