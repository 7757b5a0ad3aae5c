#!/bin/bash
# usage (under gpurun): scripts/host_chunks_ab.sh CONFIG N1 N2 ...  -- end-to-end rate (host buffers through gpmp2b_batch_optimize) by chunk count
c=$1; shift
for n in "$@"; do
  GPMP2B_HOST_CHUNKS=$n timeout 200 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-parity-sample > gpurun_out/hc.json 2> gpurun_out/hc.err
  echo "$c chunks=$n: $(python -c "
import json
d=json.load(open('gpurun_out/hc.json')); print('device %.0f traj/s (%.2f ms)  e2e %.0f traj/s' % (d['value'], d['ms_per_step'], d['e2e']['value']))" 2>&1 | tail -1)"
done
