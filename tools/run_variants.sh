#!/bin/bash
# usage (on the GPU box): bash tools/run_variants.sh "<variant names>" "<kbench shapes>" [samples]
# times the default library and each pyrism_b200/lib/variants/libprosail_<name>.so with tools/kbench.py
S=${3:-262144}
for w in $2; do
  python tools/kbench.py $S 5 $w f64
  for v in $1; do PYRISM_B200_LIB=pyrism_b200/lib/variants/libprosail_$v.so python tools/kbench.py $S 5 $w f64; done
done
