#!/bin/bash
# usage (under gpurun): scripts/mobile_ab.sh "NAME=VAL ..." ...   -- one `bench.py --config mobile` line per environment setting
for envs in "$@"; do
  tag=$(echo "$envs" | tr ' =/.' '____' | tail -c 40)
  env $envs timeout 300 python bench.py --config mobile --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/mab_$tag.json 2> gpurun_out/mab_$tag.err
  echo "$envs: $(python -c "
import json
d=json.load(open('gpurun_out/mab_$tag.json')); p=d.get('parity_sample') or {}
print('ms/step %.2f kernel_ms %.2f value %.0f e2e %.0f launches %s parity %s %.1e' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value'], d['e2e']['value'], d.get('gpu_launches'), p.get('match_frac'), p.get('max_abs_rad', 0)))" 2>&1 | tail -1)"
done
