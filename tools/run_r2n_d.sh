#!/bin/bash
# round 2n (d): packed fp32 item loop with the re-derived stages -- A/B against the round-2k library, then the fp32 tests and the sweeps
mkdir -p gpurun_out; O=gpurun_out/r2n_d.txt; : > $O
V=$PWD/pyrism_b200/lib/variants
for w in brf means four; do
  PYRISM_B200_LIB=$V/libprosail_base.so python tools/kbench.py 1000000 5 $w f32 >> $O 2>&1
  python tools/kbench.py 1000000 5 $w f32 >> $O 2>&1
done
PYRISM_B200_LIB=$V/libprosail_base.so python tools/kbench.py 8192 5 multi f32 >> $O 2>&1
python tools/kbench.py 8192 5 multi f32 >> $O 2>&1
cat $O
timeout 900 python -m pytest tests/test_gpu_fp32.py tests/test_gpu_sweeps.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -5 | tee -a $O
