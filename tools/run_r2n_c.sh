#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/r2n_c.txt; : > $O
for w in four leaf means; do
  python tools/kbench.py 1000000 5 $w f64 >> $O 2>&1
  PYRISM_B200_NO_TAIL=1 python tools/kbench.py 1000000 5 $w f64 >> $O 2>&1
  python tools/kbench.py 1000000 5 $w f64 >> $O 2>&1
  PYRISM_B200_NO_TAIL=1 python tools/kbench.py 1000000 5 $w f64 >> $O 2>&1
done
python tools/kbench.py 8192 5 multi f64 >> $O 2>&1
PYRISM_B200_NO_TAIL=1 python tools/kbench.py 8192 5 multi f64 >> $O 2>&1
cat $O
