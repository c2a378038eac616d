#!/bin/bash
# ncu evidence for the parts of the path other than the vision GEMMs: the post-process kernels of one bench step and two layers of
# the text tower.  Metric lists only (CSV on the box: full reports of these many launches exceed the copy-back limit).
mkdir -p gpurun_out
export WD_BENCH_NO_RAMP=1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size"
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fast-mode --no-torch-eager"
timeout 500 ncu --metrics $M --clock-control none -k regex:"pp_|rs_" -s 47 -c 47 --csv --log-file gpurun_out/ncu_post_r02.csv $B > gpurun_out/ncu_post_r02.log 2>&1; echo "ncu post exit $?"
timeout 500 ncu --metrics $M --clock-control none -s 330 -c 24 --csv --log-file gpurun_out/ncu_text_r02.csv python tools/text_tower_run.py lvis_v1_zh large > gpurun_out/ncu_text_r02.log 2>&1; echo "ncu text exit $?"
wc -l gpurun_out/ncu_post_r02.csv gpurun_out/ncu_text_r02.csv; tail -2 gpurun_out/ncu_text_r02.log
