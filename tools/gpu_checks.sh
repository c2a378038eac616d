#!/bin/bash
# Run the GPU parity suites file by file (each in its own process, bounded), logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in ${@:-tests/test_gpu_gemm.py tests/test_gpu_rowops.py tests/test_gpu_postprocess.py tests/test_gpu_text.py tests/test_gpu_retrieval.py tests/test_gpu_letterbox.py tests/test_gpu_mm_pipeline.py tests/test_gpu_e2e.py tests/test_gpu_facade.py tests/test_gpu_dist.py}; do
  n=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -q -s --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "== $f exit $?" | tee -a gpurun_out/summary.txt
  tail -4 gpurun_out/$n.log
done
