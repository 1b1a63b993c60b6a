#!/bin/bash
# round 2f: thread-major kc layout (128-bit loads), NU-way interleaved window reductions, NU=3 band-mean instance
mkdir -p gpurun_out; O=gpurun_out/r2f.txt; : > $O
for v in "" kc0 m3; do
  L=""; [ -n "$v" ] && L=$PWD/pyrism_b200/lib/variants/libprosail_$v.so
  for w in brf leaf means multi four; do
    S=262144; [ $w = leaf ] && S=524288; [ $w = means ] && S=1048576; [ $w = multi ] && S=8192
    PYRISM_B200_LIB=$L python tools/kbench.py $S 5 $w f64 >> $O 2>&1
  done
  PYRISM_B200_LIB=$L python tools/kbench.py 262144 5 brf f32 >> $O 2>&1
done
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 >> $O
cat $O
