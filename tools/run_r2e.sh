#!/bin/bash
# round 2e: uniform tau index, bare-soil patch after the canopy stage
mkdir -p gpurun_out; O=gpurun_out/r2e.txt; : > $O
for v in "" t0; do
  L=""; [ -n "$v" ] && L=$PWD/pyrism_b200/lib/variants/libprosail_$v.so
  for w in brf leaf means multi four; do
    S=262144; [ $w = leaf ] && S=524288; [ $w = means ] && S=1048576; [ $w = multi ] && S=8192
    PYRISM_B200_LIB=$L python tools/kbench.py $S 5 $w f64 >> $O 2>&1
  done
  PYRISM_B200_LIB=$L python tools/kbench.py 262144 5 brf f32 >> $O 2>&1
done
python tools/parity_sweep.py 8000 2>&1 | grep "float" | cut -c1-60,300- >> $O
python -m pytest tests -m gpu -q 2>&1 | tail -8 >> $O
cat $O
