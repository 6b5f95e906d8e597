#!/bin/bash
# instruction / issue counters of the range CTA kernel per variant build:
#   bash benchmarks/variants_ncu.sh name ...   (after benchmarks/variants_ab.sh ran clean)
for v in "$@"; do
  DRRT_B200_LIB=$PWD/drrt_b200/variants/$v/libdrrt_b200.so timeout 600 ncu --clock-control none \
    -k regex:range_cta_kernel -s 3 -c 1 \
    --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active \
    python benchmarks/ab.py range 2>&1 | grep -E "^\s+(gpu__|smsp__|sm__|l1tex__)" | sed "s/^/$v /"
done
