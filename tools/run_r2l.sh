#!/bin/bash
# round 2l: tensor-core LUT inversion: contract tests + throughput
mkdir -p gpurun_out; O=gpurun_out/r2l.txt; : > $O
timeout 300 python -m pytest tests/test_gpu_inversion.py -m gpu -q -x 2>&1 | tail -15 >> $O
timeout 120 python tools/invbench.py 1000000 16384 15 >> $O 2>&1
timeout 120 python tools/invbench.py 12500000 16384 15 >> $O 2>&1
PYRISM_B200_INVERT_TC=0 timeout 120 python tools/invbench.py 1000000 16384 15 >> $O 2>&1
cat $O
