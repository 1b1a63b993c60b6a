#!/bin/bash
# round 2b: benign-site second-order Newton steps + fp32 tail of the tau polynomial: speed (A/B variants) and parity
mkdir -p gpurun_out; O=gpurun_out/r2b.txt; : > $O
for v in "" q1t0 q0t1 q0t0; do
  L=""; [ -n "$v" ] && L=$PWD/pyrism_b200/lib/variants/libprosail_$v.so
  for w in brf leaf; do
    S=262144; [ $w = leaf ] && S=524288
    PYRISM_B200_LIB=$L python tools/kbench.py $S 5 $w f64 >> $O 2>&1
  done
done
python tools/parity_sweep.py 8000 >> $O 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $O
cat $O
