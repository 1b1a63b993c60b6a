#!/bin/bash
# On the GPU box: tests, bench (no profiler), then the ncu launch list of the same bench command and one full capture of the
# dominant kernel in fp64 and in fp32.  Results under gpurun_out/.
set -x
R=${1:-r1b}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/gpu_tests_$R.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.log 2> gpurun_out/bench_$R.err || exit 1
tail -1 gpurun_out/bench_$R.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 2 --warmup 1 --samples 262144 --no-cpu --no-e2e > gpurun_out/ncu_list_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bands -s 3 -c 1 -o gpurun_out/prof_${R}_f64 -f \
    python tools/kbench.py 262144 3 brf f64 > gpurun_out/ncu_f64_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bands -s 3 -c 1 -o gpurun_out/prof_${R}_f32 -f \
    python tools/kbench.py 262144 3 brf f32 > gpurun_out/ncu_f32_$R.log 2>&1
cat gpurun_out/gpu_tests_$R.log
ls -la gpurun_out/
