#!/bin/bash
# Round-2 ncu evidence: launch list of one bench step (shares), full captures of the dominant kernels with source.
mkdir -p gpurun_out
export WD_BENCH_NO_RAMP=1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fast-mode --no-torch-eager"
# warm-up: 3 eager test_step + ... the last graph replay is what we want: list every launch, post-filter here
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_list_r02.log 2>&1; echo "ncu list exit $?"
wc -l gpurun_out/launches_r02.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_split -s 712 -c 6 -o gpurun_out/prof_split_r02 $B > gpurun_out/ncu_split_r02.log 2>&1; echo "ncu split exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv7_tma -s 150 -c 2 -o gpurun_out/prof_dw_r02 $B > gpurun_out/ncu_dw_r02.log 2>&1; echo "ncu dw exit $?"
timeout 600 ncu --set full --clock-control none -k regex:ln_rows -s 166 -c 2 -o gpurun_out/prof_ln_r02 $B > gpurun_out/ncu_ln_r02.log 2>&1; echo "ncu ln exit $?"
ls -la gpurun_out/*.ncu-rep | tail -5
