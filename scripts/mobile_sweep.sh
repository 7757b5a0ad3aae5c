#!/bin/bash
# usage (under gpurun): scripts/mobile_sweep.sh B1 B2 ...  -- one-kernel optimizer vs pipeline for config 4 at several batch sizes
for B in "$@"; do for lie in 0 1; do
  GPMP2B_PK_MIN_BATCH=1 GPMP2B_PK_LIE=$lie timeout 300 python bench.py --config mobile --batch $B --steps 6 --warmup 3 --no-cpu-baseline --no-parity-sample > gpurun_out/msw.json 2> gpurun_out/msw.err
  echo "B=$B pipeline=$lie: $(python -c "
import json
d=json.load(open('gpurun_out/msw.json'))
print('ms/step %.3f value %.0f e2e %.0f launches %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d.get('gpu_launches')))" 2>&1 | tail -1)"
done; done
