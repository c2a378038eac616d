#!/bin/bash
mkdir -p gpurun_out
python tools/split_sweep.py gpurun_out/split_sweep_q.json 2>&1 | grep -E "s2 pw1 gelu|s2 pw2 resid-inplace|s0 pw1 gelu|s1 pw1|neck" 
bash tools/gpu_quick.sh
