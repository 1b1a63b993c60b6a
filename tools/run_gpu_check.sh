set -x
python -m pytest tests/test_gpu_fp32.py -x -q 2>&1 | tail -25 > gpurun_out/fp32_tests.log
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fp32.py 2>&1 | tail -8 > gpurun_out/gpu_tests.log
for w in brf four leaf means; do for d in f64 f32; do python tools/kbench.py 262144 5 $w $d; done; done > gpurun_out/kbench_r1b.log 2>&1
cat gpurun_out/fp32_tests.log gpurun_out/gpu_tests.log gpurun_out/kbench_r1b.log
