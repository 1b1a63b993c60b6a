#!/bin/bash
# round 2d: full GPU test suite, bench lines of every config on one GPU, ncu capture of the fp64 BRF kernel, D2H probe
mkdir -p gpurun_out; O=gpurun_out/r2d.txt; : > $O
python -m pytest tests -m gpu -q 2>&1 | tail -6 >> $O
python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench_c2.json 2>> $O
for c in 1 3 4 0; do python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r2d_bench_c$c.json 2>> $O; done
tools/micro/bin/d2h_probe 1 1024 6 > gpurun_out/r2d_d2h_probe_1gpu.txt 2>&1
python tools/kbench.py 262144 3 brf f64 >> $O 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_bands -s 3 -c 1 -f -o gpurun_out/prof_r2d_f64 python tools/kbench.py 262144 3 brf f64 > gpurun_out/r2d_ncu.log 2>&1
cat $O; for c in 2 1 3 4 0; do cut -c1-1500 gpurun_out/r2d_bench_c$c.json; echo; done; cat gpurun_out/r2d_d2h_probe_1gpu.txt; tail -3 gpurun_out/r2d_ncu.log
