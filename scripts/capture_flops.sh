#!/bin/bash
# usage (under gpurun): scripts/capture_flops.sh TAG  -- executed FP64 work of round 0 of the pipeline (every trajectory
# linearizes, solves and evaluates once): DFMA / DMUL / DADD thread instructions and DMMA warp instructions per kernel,
# next to the algorithmic counts of SURVEY.md 8(d)  ->  gpurun_out/TAG_executed_flops.txt
tag=$1
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed_pipe_tensor_subpipe_dmma.sum,smsp__inst_executed.sum,gpu__time_duration.sum
for c in wam mobile; do
  ncu --metrics $M --clock-control none -k regex:pk_ -s 1 -c 3 --csv --log-file gpurun_out/${tag}_flops_$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-parity-sample > /dev/null 2>&1
done
python scripts/executed_flops.py gpurun_out/${tag}_flops_wam.csv wam gpurun_out/${tag}_flops_mobile.csv mobile > gpurun_out/${tag}_executed_flops.txt 2>&1
cat gpurun_out/${tag}_executed_flops.txt
