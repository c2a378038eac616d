#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/split_sweep.py gpurun_out/split_sweep.json 2>&1 | tail -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_split -s 3 -c 1 -o gpurun_out/prof_split_pw1 python tools/split_sweep.py - "s2 pw1 gelu" > gpurun_out/ncu_split_pw1.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/ncu_split_pw1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_split -s 3 -c 1 -o gpurun_out/prof_split_pw2 python tools/split_sweep.py - "s2 pw2 res" > gpurun_out/ncu_split_pw2.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep | tail -3
