#!/bin/bash
# usage (under gpurun): scripts/pk_bench.sh variant ...   -- bit-identity against the fused kernel, then the bench line
export PK_AB_WAM_ONLY=1
GPMP2B_PK=0 python scripts/pk_ab.py gpurun_out/pk_base.npz > gpurun_out/pk_base.log 2>&1
for v in "$@"; do
  export GPMP2B_LIB=$PWD/variants/lib_$v.so GPMP2B_PK=1
  timeout 300 python scripts/pk_ab.py gpurun_out/pk_$v.npz > gpurun_out/pk_$v.log 2>&1
  c=$(python scripts/pk_ab.py compare gpurun_out/pk_base.npz gpurun_out/pk_$v.npz 2>&1 | tail -1)
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity-sample > gpurun_out/pkb_$v.json 2> gpurun_out/pkb_$v.err
  echo "$v: [$c] $(python -c "import json; d=json.load(open('gpurun_out/pkb_$v.json')); print('ms/step %.2f kernel_ms %.2f value %.0f e2e %.0f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value'], d['e2e']['value']))" 2>&1 | tail -1)"
  rm -f gpurun_out/pk_$v.npz
done
rm -f gpurun_out/pk_base.npz
