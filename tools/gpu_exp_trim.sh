#!/bin/bash
mkdir -p gpurun_out
python tools/split_sweep.py gpurun_out/split_sweep_trim.json 2>&1 | tail -24
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -3
timeout 420 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-fast-mode --no-torch-eager --profile-ops gpurun_out/r2_ops_trim.json > gpurun_out/r2_bench_trim.json 2> gpurun_out/r2_bench_trim.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_trim.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'])
p=json.load(open('gpurun_out/r2_ops_trim.json'))
for k,v in p['families'].items(): print(k, round(v['ms'],3), v['launches'], v['tflops'] and round(v['tflops'],1))
"; tail -3 gpurun_out/r2_bench_trim.err
