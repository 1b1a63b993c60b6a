#!/bin/bash
# round 2n: A/B of the FP64-instruction reductions (base = the round-2m library) + the gpu test suite on the new library
# (base: `git worktree add /tmp/base 5859ab0 && make -C /tmp/base && cp /tmp/base/pyrism_b200/lib/libprosail_b200.so pyrism_b200/lib/variants/libprosail_base.so`;
#  the variants directory is git-ignored and was removed again after the round)
mkdir -p gpurun_out; O=gpurun_out/r2n.txt; : > $O
V=$PWD/pyrism_b200/lib/variants
for w in "brf f64" "leaf f64" "means f64" "four f64" "brf f32"; do
  set -- $w
  for lib in base ""; do
    if [ -n "$lib" ]; then export PYRISM_B200_LIB=$V/libprosail_$lib.so; else unset PYRISM_B200_LIB; fi
    python tools/kbench.py 1000000 5 $1 $2 >> $O 2>&1
  done
done
for lib in base ""; do
  if [ -n "$lib" ]; then export PYRISM_B200_LIB=$V/libprosail_$lib.so; else unset PYRISM_B200_LIB; fi
  python tools/kbench.py 8192 5 multi f64 >> $O 2>&1
done
unset PYRISM_B200_LIB
for x in $EXTRA_VARIANTS; do PYRISM_B200_LIB=$V/libprosail_$x.so python tools/kbench.py 1000000 5 brf f64 >> $O 2>&1; done
cat $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee -a $O
