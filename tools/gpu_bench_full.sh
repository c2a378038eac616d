#!/bin/bash
# the driver's bench invocations + the side workloads; $1 = tag
T=${1:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --profile-ops gpurun_out/ops_profile_$T.json > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
echo "bench exit $?"; tail -c 6000 gpurun_out/bench_$T.json; tail -5 gpurun_out/bench_$T.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err
echo "ref exit $?"; tail -c 1200 gpurun_out/bench_ref_$T.json
for WL in uni_proposals corpus; do
  timeout 600 python bench.py --workload $WL --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_${WL}_$T.json 2> gpurun_out/bench_${WL}_$T.err
  echo "$WL exit $?"; tail -c 1800 gpurun_out/bench_${WL}_$T.json; tail -3 gpurun_out/bench_${WL}_$T.err
done
