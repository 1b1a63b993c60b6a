#!/bin/bash
# round 2n (b): configuration variants on the new code
mkdir -p gpurun_out; O=gpurun_out/r2n_b.txt; : > $O
V=$PWD/pyrism_b200/lib/variants
run() { lib=$1; shift; if [ -n "$lib" ]; then export PYRISM_B200_LIB=$V/libprosail_$lib.so; else unset PYRISM_B200_LIB; fi; python tools/kbench.py "$@" >> $O 2>&1; }
for lib in "" nonu3 chunk16; do run "$lib" 1000000 5 brf f64; done
for lib in "" nonu3 chunk16; do run "$lib" 1000000 5 leaf f64; done
for lib in "" four3; do run "$lib" 1000000 5 four f64; done
for lib in "" means3; do run "$lib" 1000000 5 means f64; done
for lib in "" f32w24 f32w20 f32w12x2; do run "$lib" 1000000 5 brf f32; done
for lib in "" f32w24 f32w20 f32w12x2; do run "$lib" 1000000 5 means f32; done
cat $O
