#!/bin/bash
# A/B the variants built by scripts/build_variant_lie.sh on the GPU box: mobile parity tests, then the config-4 bench.
# usage (under gpurun): scripts/ab_lie.sh name1 name2 ...   ("base" = the in-tree library)
for v in "$@"; do
  if [ "$v" = base ]; then unset GPMP2B_LIB; else export GPMP2B_LIB=$PWD/variants/lib_$v.so; fi
  t=$(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mobile_optimize or mobile_linearize" 2>&1 | tail -1)
  python bench.py --config mobile --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/abl_$v.json 2> gpurun_out/abl_$v.err
  echo "$v: tests[$t] $(python -c "import json,sys; d=json.load(open('gpurun_out/abl_$v.json')); p=d.get('parity_sample') or {}; print('ms/step %.2f kernel_ms %.2f value %.0f e2e %.0f parity %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value'], d['e2e']['value'], p.get('match_frac')))" 2>&1 | tail -1)"
done
