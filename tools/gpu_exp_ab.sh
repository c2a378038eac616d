#!/bin/bash
# A/B of an experimental library variant against the tree's library on the sweep shapes; $1 = variant .so
for c in "s0 pw1 gelu f16x2" "s1 pw1" "s2 pw1 gelu" "neck silu f16x2 N128"; do
  echo "--- base"; python tools/split_sweep.py - "$c" 2>&1 | tail -1
  echo "--- variant"; WD_LIB_PATH=$PWD/$1 python tools/split_sweep.py - "$c" 2>&1 | tail -1
done
