#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/r2m.txt; : > $O
timeout 240 python -m pytest tests/test_gpu_inversion.py -m gpu -q -x 2>&1 | tail -5 >> $O
timeout 120 python tools/invbench.py 1000000 16384 15 >> $O 2>&1
timeout 120 python tools/invbench.py 12500000 16384 15 >> $O 2>&1
timeout 120 python tools/invbench.py 12500000 131072 15 >> $O 2>&1
cat $O
