#!/bin/bash
# ncu --set full captures: a stage-2 depthwise launch, the first stage-0 pw1 GEMM (K = 128, GELU epilogue bound), letterbox + retrieval kernels
mkdir -p gpurun_out
T=${1:-r01c}
export WD_BENCH_NO_RAMP=1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-mode"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv7_tma -s 10 -c 1 -o gpurun_out/prof_dw_$T $B > gpurun_out/ncu_dw_$T.log 2>&1; echo "ncu dw exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc_kernel<256, __nv_bfloat16" -s 0 -c 1 -o gpurun_out/prof_gemm_s0pw1_$T $B > gpurun_out/ncu_gemm_s0_$T.log 2>&1; echo "ncu gemm s0 exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:lb_pass|retr_reduce|scale_rows" -s 4 -c 6 -o gpurun_out/prof_aux_$T python tools/bench_aux.py > gpurun_out/ncu_aux_$T.log 2>&1; echo "ncu aux exit $?"
ls -la gpurun_out/*.ncu-rep
