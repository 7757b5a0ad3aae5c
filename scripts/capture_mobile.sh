#!/bin/bash
# usage (under gpurun): scripts/capture_mobile.sh TAG   -- config 4 (Pose2MobileArm) on the phase-kernel pipeline: bench line,
# ncu launch list with DRAM bytes, ncu --set full summaries of its three kernels (reports summarised on the box and deleted)
tag=$1
out=gpurun_out
python bench.py --config mobile --steps 10 --warmup 3 > $out/${tag}_bench_mobile.json 2> $out/${tag}_bench_mobile.err
B="python bench.py --config mobile --steps 1 --warmup 1 --no-cpu-baseline --no-parity-sample"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 0 -c 700 --csv --log-file $out/${tag}_launches_mobile.csv $B > /dev/null 2>&1
cap() {  # name kernel_regex skip
  ncu --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 -o /tmp/${tag}_$1 -f $B > /dev/null 2>&1
  python scripts/ncu_summary.py /tmp/${tag}_$1.ncu-rep $2 "ncu --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 $B" > $out/${tag}_ncu_mobile_$1.txt 2>&1
  rm -f /tmp/${tag}_$1.ncu-rep
}
cap lin pk_lin_full 4
cap solve pk_solve_mma_h 6
cap err pk_err 8
ls -la $out/${tag}_*
