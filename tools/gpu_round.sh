#!/bin/bash
# One GPU visit: parity suites, smoke, side measurements, bench line.  $1 = tag for the output names.
T=${1:-run}
mkdir -p gpurun_out
bash tools/gpu_checks.sh
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$T.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$T.log
timeout 400 python tools/bench_aux.py > gpurun_out/bench_aux_$T.json 2> gpurun_out/bench_aux_$T.err; echo "aux exit $?"; tail -3 gpurun_out/bench_aux_$T.err
timeout 420 python bench.py --gpus 1 --steps 10 --warmup 3 --profile-ops gpurun_out/ops_profile_$T.json > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench_$T.json; tail -3 gpurun_out/bench_$T.err
