#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/r2_gemm.log 2>&1
echo "== gemm exit $?"; tail -3 gpurun_out/r2_gemm.log
timeout 600 python tools/split_sweep.py gpurun_out/split_sweep.json 2>&1 | grep -v "lblk=2" | awk 'NR%2==1' | tail -12
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -s --timeout 800 --timeout-method=thread -p no:cacheprovider -k "north_star" > gpurun_out/r2_e2e.log 2>&1
echo "== e2e exit $?"; grep -E "_precise_" gpurun_out/r2_e2e.log | python -c "
import sys,json
for line in sys.stdin:
    j=line[line.index('{'):]; d=json.loads(j)
    print(line[:line.index('{')], {k:v[0] for k,v in d.items() if k.startswith(('p5','logit','dist'))})
"; tail -3 gpurun_out/r2_e2e.log
timeout 420 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-fast-mode --no-torch-eager --profile-ops gpurun_out/r2_ops.json > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "== bench exit $?"; python -c "
import json;d=json.load(open('gpurun_out/r2_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'])
p=json.load(open('gpurun_out/r2_ops.json'))
for k,v in p['families'].items(): print(k, round(v['ms'],3), v['launches'], v['tflops'] and round(v['tflops'],1))
"; tail -3 gpurun_out/r2_bench.err
