#!/bin/bash
# round 2i: disjoint-piece segmented reduction of the band means; nested-FMA tails of the diffuse / coupling stages
mkdir -p gpurun_out; O=gpurun_out/r2i.txt; : > $O
for v in "" nored; do
  L=""; [ -n "$v" ] && L=$PWD/pyrism_b200/lib/variants/libprosail_$v.so
  PYRISM_B200_LIB=$L python tools/kbench.py 1048576 5 means f64 >> $O 2>&1
done
for w in brf leaf multi four; do
  S=262144; [ $w = leaf ] && S=524288; [ $w = multi ] && S=8192
  python tools/kbench.py $S 5 $w f64 >> $O 2>&1
done
python tools/kbench.py 262144 5 brf f32 >> $O 2>&1
python tools/kbench.py 1048576 5 means f32 >> $O 2>&1
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 >> $O
cat $O
