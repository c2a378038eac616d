#!/bin/bash
mkdir -p gpurun_out
python tools/split_sweep.py gpurun_out/split_sweep_lblk.json 2>&1 | grep -v "bn=128 lblk=1" | tail -24 | awk '{ if (NR%3!=0) print }'
for L in 1 2; do
  echo "#### WD_SPLIT_LBLK=$L"
  WD_SPLIT_LBLK=$L bash tools/gpu_e2e.sh "north_star or benched" 2>&1 | grep -v "^\s*$" | cut -c1-330 | tail -22
  WD_SPLIT_LBLK=$L timeout 420 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-fast-mode --no-torch-eager --profile-ops gpurun_out/r2_ops_l$L.json > gpurun_out/r2_bench_l$L.json 2> gpurun_out/r2_bench_l$L.err
  echo "== bench lblk=$L exit $?"; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_l$L.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'])
p=json.load(open('gpurun_out/r2_ops_l$L.json'))
for k,v in p['families'].items(): print(k, round(v['ms'],3), v['launches'], v['tflops'] and round(v['tflops'],1))
"; tail -3 gpurun_out/r2_bench_l$L.err
done
