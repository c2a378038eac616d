#!/bin/bash
# ncu captures of the hot kernels inside one bench step (1 GPU).  $1 = tag
mkdir -p gpurun_out
T=${1:-r01}
export WD_BENCH_NO_RAMP=1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-mode"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 330 --csv --log-file gpurun_out/launches_$T.csv $B > gpurun_out/ncu_list_$T.log 2>&1; echo "ncu list exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 20 -c 2 -o gpurun_out/prof_gemm_$T $B > gpurun_out/ncu_gemm_$T.log 2>&1; echo "ncu gemm exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv7_tma -s 10 -c 1 -o gpurun_out/prof_dw_$T $B > gpurun_out/ncu_dw_$T.log 2>&1; echo "ncu dw exit $?"
ls -la gpurun_out | tail -6
