#!/bin/bash
# usage (under gpurun): scripts/capture_final.sh TAG   -- the evidence set of a revision in one call: bench lines of every
# BASELINE config and of the reference arm, ncu launch lists with DRAM bytes (WAM and config 4), ncu --set full summaries
# of the kernels on the benched paths (reports summarised on the box and deleted: gpurun brings back 64 MiB)
tag=$1
out=gpurun_out
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>/dev/null
for c in planar2 planar3gp mobile; do python bench.py --config $c --steps 5 --warmup 3 > $out/${tag}_bench_$c.json 2>/dev/null; done
B="--steps 1 --warmup 1 --no-cpu-baseline --no-parity-sample"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 0 -c 700 --csv --log-file $out/${tag}_launches.csv python bench.py $B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 0 -c 700 --csv --log-file $out/${tag}_launches_mobile.csv python bench.py --config mobile $B > /dev/null 2>&1
cap() {  # name kernel_regex skip "bench args"
  local cmd="python bench.py $4 $B"
  ncu --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 -o /tmp/${tag}_$1 -f $cmd > /dev/null 2>&1
  python scripts/ncu_summary.py /tmp/${tag}_$1.ncu-rep $2 "ncu --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 $cmd" > $out/${tag}_ncu_$1.txt 2>&1
  rm -f /tmp/${tag}_$1.ncu-rep
}
cap lin pk_linh 4 ""
cap solve pk_solve_mma_h 6 ""
cap err pk_err 8 ""
cap mobile_lin pk_lin_full 4 "--config mobile"
cap mobile_solve pk_solve_mma_h 6 "--config mobile"
cap mobile_err pk_err 8 "--config mobile"
ls -la $out/${tag}_*
