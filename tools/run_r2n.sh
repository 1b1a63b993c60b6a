#!/bin/bash
# round 2n: final code -- smoke, full GPU test suite, bench lines
mkdir -p gpurun_out; O=gpurun_out/r2n.txt; : > $O
python __graft_entry__.py smoke >> $O 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -4 >> $O
python bench.py > gpurun_out/r2n_bench_c2.json 2>> $O
python bench.py --config 4 --steps 5 > gpurun_out/r2n_bench_c4.json 2>> $O
cat $O; tail -c 1500 gpurun_out/r2n_bench_c4.json
