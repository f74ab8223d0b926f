#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): plain run first, then the launch list, then one full capture
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -s 24400 -c 1700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 11700 -c 4 -o gpurun_out/prof_gemm_tc_v4 $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1930 -c 3 -o gpurun_out/prof_attention_tc_v4 $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
