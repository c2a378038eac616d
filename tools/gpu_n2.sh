#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 500 -p no:cacheprovider 2>&1 | tail -3
for WL in detect uni_proposals corpus; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --workload $WL > gpurun_out/bench_n2_${WL}_r02.json 2> gpurun_out/bench_n2_${WL}_r02.err
  echo "$WL n2 exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_n2_${WL}_r02.json'));print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['clocks'])"; tail -2 gpurun_out/bench_n2_${WL}_r02.err
done
