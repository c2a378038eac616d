#!/bin/bash
# Round-2 experiment 1: the fp16 hi/lo GEMM kernel.  Parity of the op tests, e2e logit error as a function of the TMEM
# accumulator block length (WD_SPLIT_LBLK), and a first timing of the default mode.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q -s -x --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/r2_gemm.log 2>&1
echo "== gemm exit $?"; tail -5 gpurun_out/r2_gemm.log
timeout 600 python -m pytest tests/test_gpu_rowops.py -m gpu -q -s --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/r2_rowops.log 2>&1
echo "== rowops exit $?"; tail -5 gpurun_out/r2_rowops.log
for L in 1 2 4 1000; do
  WD_SPLIT_LBLK=$L timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q -s --timeout 500 --timeout-method=thread -p no:cacheprovider \
     -k "north_star and (base-80 or tiny-5)" > gpurun_out/r2_e2e_lblk$L.log 2>&1
  echo "== e2e lblk=$L exit $?"; grep -E "^(tiny|base)_" gpurun_out/r2_e2e_lblk$L.log | cut -c1-900; tail -2 gpurun_out/r2_e2e_lblk$L.log
done
for L in 1 2; do
  WD_SPLIT_LBLK=$L timeout 420 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-parity-mode --profile-ops gpurun_out/r2_ops_lblk$L.json > gpurun_out/r2_bench_lblk$L.json 2> gpurun_out/r2_bench_lblk$L.err
  echo "== bench lblk=$L exit $?"; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_lblk$L.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'])
p=json.load(open('gpurun_out/r2_ops_lblk$L.json'))
for k,v in p['families'].items(): print(k, round(v['ms'],3), v['launches'], v['tflops'] and round(v['tflops'],1))
"; tail -3 gpurun_out/r2_bench_lblk$L.err
done
