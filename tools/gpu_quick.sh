#!/bin/bash
# quick regression + bench line after a kernel change
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -2
bash tools/gpu_e2e.sh "north_star or benched" 2>&1 | grep -E "passed|failed|FAILED" | tail -3
timeout 420 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-fast-mode --no-torch-eager --profile-ops gpurun_out/r2_ops_q.json > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_q.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'])
p=json.load(open('gpurun_out/r2_ops_q.json'))
for k,v in p['families'].items(): print(k, round(v['ms'],3), v['launches'], v['tflops'] and round(v['tflops'],1))
"; tail -3 gpurun_out/r2_bench_q.err
