#!/bin/bash
mkdir -p gpurun_out
python tools/split_sweep.py gpurun_out/split_sweep_s0.json "s0 " 2>&1 | tail -12
python tools/split_sweep.py - "s1 " 2>&1 | tail -3
for c in "s0 pw1 gelu f16x2" "s0 pw2 inplace"; do
  n=$(echo "$c" | tr ' ' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_split --launch-skip 5 --launch-count 1 -f -o gpurun_out/prof_$n python tools/split_sweep.py - "$c" > gpurun_out/ncu_$n.log 2>&1
  echo "ncu $c exit $?"
done
ls -la gpurun_out/*.ncu-rep
