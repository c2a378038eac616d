#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_e2e.py -m gpu -q -s --timeout 1200 --timeout-method=thread -p no:cacheprovider ${1:+-k "$1"} > gpurun_out/r2_e2e.log 2>&1
echo "== e2e exit $?"; grep -E "_(precise|fast)_" gpurun_out/r2_e2e.log | python -c "
import sys,json
for line in sys.stdin:
    if '{' not in line: continue
    j=line[line.index('{'):]; d=json.loads(j)
    print(line[:line.index('{')], {k:(v[0] if isinstance(v,list) else v) for k,v in d.items() if k.startswith(('p5','logit','dist')) and 'fp64' not in k})
    for k,v in d.items():
        if 'fp64' in k: print('    ',k,v)
"; grep -E "max_over_tol|passed|failed|Error" gpurun_out/r2_e2e.log | cut -c1-300 | tail -12
