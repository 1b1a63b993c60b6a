#!/bin/bash
# round 2h: window map hoisted out of the item loop (band-mean instance); what the reductions cost (nored = no butterflies)
mkdir -p gpurun_out; O=gpurun_out/r2h.txt; : > $O
for v in "" nored; do
  L=""; [ -n "$v" ] && L=$PWD/pyrism_b200/lib/variants/libprosail_$v.so
  PYRISM_B200_LIB=$L python tools/kbench.py 1048576 5 means f64 >> $O 2>&1
  PYRISM_B200_LIB=$L python tools/kbench.py 1048576 5 means f32 >> $O 2>&1
done
python tools/kbench.py 262144 5 brf f64 >> $O 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_inversion.py tests/test_gpu_fp32.py -m gpu -q -x 2>&1 | tail -5 >> $O
cat $O
