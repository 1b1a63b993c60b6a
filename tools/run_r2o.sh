#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/r2o.txt; : > $O
timeout 400 python -m pytest tests/test_gpu_inversion.py -m gpu -q -x 2>&1 | tail -15 >> $O
for K in 8e-6 2e-6 1e-6 5e-7; do
  PYRISM_B200_TC_KAPPA=$K timeout 200 python tools/invbench.py 12500000 16384 15 real >> $O 2>&1
done
timeout 120 python tools/invbench.py 12500000 16384 15 random >> $O 2>&1
PYRISM_B200_INVERT_TC=0 timeout 200 python tools/invbench.py 12500000 16384 15 real >> $O 2>&1
cat $O
