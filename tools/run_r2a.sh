#!/bin/bash
# round 2a: pipe microbenchmark for conversions / seeds + today's baselines of the kernels
mkdir -p gpurun_out
tools/micro/bin/conv_mix > gpurun_out/r2a_conv_mix.txt 2>&1
for w in brf leaf means four multi; do
  S=262144; [ $w = multi ] && S=8192; [ $w = leaf ] && S=524288; [ $w = means ] && S=1048576
  python tools/kbench.py $S 5 $w f64 >> gpurun_out/r2a_kbench.txt 2>&1
done
python tools/kbench.py 262144 5 brf f32 >> gpurun_out/r2a_kbench.txt 2>&1
python tools/kbench.py 1000000 5 brf f64 >> gpurun_out/r2a_kbench.txt 2>&1
cat gpurun_out/r2a_conv_mix.txt gpurun_out/r2a_kbench.txt
