#!/bin/bash
# round 2n: final-code bench lines (all configs, 1 GPU), ncu launch list of the bench command, ncu --set full captures of the
# kernel instances (summarised ON the box by tools/ncu_summary.py: only the text summaries and one report travel back)
mkdir -p gpurun_out; O=gpurun_out/r2n_final.txt; : > $O
python bench.py --steps 10 --warmup 3 > gpurun_out/r2n_bench_c2_n1.json 2>> $O
for c in 1 3 4; do python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r2n_bench_c${c}_n1.json 2>> $O; done
python bench.py --config 0 --steps 200 --warmup 20 > gpurun_out/r2n_bench_c0_n1.json 2>> $O
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_bench_reference_n1.json 2>> $O
python bench.py --steps 2 --warmup 1 --samples 262144 --no-cpu --no-e2e > gpurun_out/r2n_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches_bench_262144.csv \
    python bench.py --steps 2 --warmup 1 --samples 262144 --no-cpu --no-e2e > gpurun_out/r2n_ncu_list.log 2>&1
for w in "brf f64 22" "leaf f64 22" "brf f32 33" "means f64 13" "multi f64 33" "four f64 33"; do
  set -- $w; S=262144; [ $1 = multi ] && S=4096
  ITEMS=$((S * $3)); [ $1 = multi ] && ITEMS=$((S * 64 * $3))
  python tools/kbench.py $S 3 $1 $2 >> $O 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_bands -s 3 -c 1 -f -o /tmp/prof_r2n_$1_$2 python tools/kbench.py $S 3 $1 $2 > gpurun_out/r2n_ncu_$1_$2.log 2>&1
  python tools/ncu_summary.py /tmp/prof_r2n_$1_$2.ncu-rep gpurun_out/r2n_k_bands_$1_$2.txt $ITEMS > /dev/null 2>> $O
done
cp /tmp/prof_r2n_brf_f64.ncu-rep gpurun_out/
# memory checker on a tiny pass through every kernel and entry point (closed on this pool in round 1: the attempt is recorded either way)
(compute-sanitizer --version; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/all_paths.py) > gpurun_out/r2n_sanitizer.txt 2>&1; tail -15 gpurun_out/r2n_sanitizer.txt >> $O
cat $O; ls -la gpurun_out/; du -sh gpurun_out
