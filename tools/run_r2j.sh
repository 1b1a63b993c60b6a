#!/bin/bash
# round 2j: Ap / Am form of the leaf-dependent canopy factors (one FMA per u1,u2,v1,v2), cancellation-free m^2, Stokes front end
mkdir -p gpurun_out; O=gpurun_out/r2j.txt; : > $O
for w in brf leaf multi four means; do
  S=262144; [ $w = leaf ] && S=524288; [ $w = multi ] && S=8192;  [ $w = means ] && S=1048576
  python tools/kbench.py $S 5 $w f64 >> $O 2>&1
done
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 >> $O
cat $O
