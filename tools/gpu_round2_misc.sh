#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rowops.py tests/test_gpu_facade.py tests/test_gpu_text.py -m gpu -q --timeout 800 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --size tiny --batch 1 --classes 5 --steps 50 --warmup 5 --no-torch-eager > gpurun_out/bench_c1_r02.json 2> gpurun_out/bench_c1_r02.err; echo "c1 exit $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_c1_r02.json'));print('C1', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['workload'], d.get('fast_mode',{}).get('value'), d['cpu_baseline'])"
timeout 900 python bench.py --size large --batch 16 --res 800 --classes 1203 --steps 5 --warmup 3 --no-torch-eager --no-cpu-baseline > gpurun_out/bench_c3_r02.json 2> gpurun_out/bench_c3_r02.err; echo "c3 exit $?"; tail -2 gpurun_out/bench_c3_r02.err
python -c "
import json;d=json.load(open('gpurun_out/bench_c3_r02.json'));print('C3', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['workload'], d.get('fast_mode',{}).get('value'), d['roofline']['all_gemm'])"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -p no:cacheprovider -k "split_precise or (split_shapes and 515) or (split_shapes and 38400-192) or (inplace_residual and 515) or dfl_epilogue_split or (conv3x3_split_shapes and 16-16)" > gpurun_out/sanitizer_racecheck_r02.log 2>&1; echo "racecheck exit $?"; tail -6 gpurun_out/sanitizer_racecheck_r02.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -p no:cacheprovider -k "split_precise or (split_shapes and 515) or (split_shapes and 38400-192) or inplace_residual or dfl_epilogue_split or conv3x3_split_shapes" > gpurun_out/sanitizer_memcheck_r02.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/sanitizer_memcheck_r02.log
for WL in uni_proposals corpus; do
  timeout 600 python bench.py --workload $WL --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_${WL}_r02c.json 2> gpurun_out/bench_${WL}_r02c.err
  echo "$WL exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_${WL}_r02c.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
