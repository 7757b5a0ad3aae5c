#!/bin/bash
# A/B the variants built by scripts/build_variant.sh on the GPU box: WAM parity tests, then the default bench.
# usage (under gpurun): scripts/ab.sh name1 name2 ...   ("base" = the in-tree library)
for v in "$@"; do
  if [ "$v" = base ]; then unset GPMP2B_LIB; else export GPMP2B_LIB=$PWD/variants/lib_$v.so; fi
  t=$(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wam or WAM" 2>&1 | tail -1)
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  echo "$v: tests[$t] $(python -c "import json,sys; d=json.load(open('gpurun_out/ab_$v.json')); print('ms/step %.2f kernel_ms %.2f value %.0f e2e %.0f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value'], d['e2e']['value']))" 2>&1 | tail -1)"
done
