#!/bin/bash
# BASELINE configs[3] / [4] on the 8 GPUs of one box (north_star's multi-GPU shapes): one rank per GPU over NCCL
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10
for WL in uni_proposals corpus; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 10 --warmup 3 --workload $WL > gpurun_out/bench_n8_${WL}_r02.json 2> gpurun_out/bench_n8_${WL}_r02.err
  echo "$WL n8 exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_n8_${WL}_r02.json'));print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['config'].get('exchange'), d['clocks'])"; tail -2 gpurun_out/bench_n8_${WL}_r02.err
done
