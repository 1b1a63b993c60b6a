#!/bin/bash
# round 2c: SFU / integer instructions beside a busy FP64 pipe (micro), merged-reciprocal build with second-order benign steps
mkdir -p gpurun_out; O=gpurun_out/r2c.txt; : > $O
tools/micro/bin/pipe_share > gpurun_out/r2c_pipe_share.txt 2>&1
for v in "" t0 nu2 q0t0; do
  L=""; [ -n "$v" ] && L=$PWD/pyrism_b200/lib/variants/libprosail_$v.so
  for w in brf leaf means multi; do
    S=262144; [ $w = leaf ] && S=524288; [ $w = means ] && S=1048576; [ $w = multi ] && S=8192
    PYRISM_B200_LIB=$L python tools/kbench.py $S 5 $w f64 >> $O 2>&1
  done
done
python tools/parity_sweep.py 8000 2>&1 | grep float64 >> $O
python -m pytest tests -m gpu -q 2>&1 | tail -15 >> $O
cat gpurun_out/r2c_pipe_share.txt $O
