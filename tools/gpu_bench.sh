#!/bin/bash
# First measurement pass: bench line + per-op table + ncu launch list + one full ncu capture of the top GEMM.
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 420 python bench.py --gpus 1 --steps 10 --warmup 3 --profile-ops gpurun_out/ops_profile.json > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 330 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_bench.log 2>&1
  echo "ncu list exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 3 -o gpurun_out/prof_gemm \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
