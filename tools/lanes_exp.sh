#!/bin/bash
# perf experiment: extraction + match step time vs number of concurrent frame-range lanes
for l in 1 2 3 4; do
  ORBX_LANES=$l timeout 300 python bench.py --steps 100 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes=$l', round(d['value']), 'fps', round(d['ms_per_step'],3), 'ms/step')"
done
