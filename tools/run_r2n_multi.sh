#!/bin/bash
# round 2n (8-GPU box): the final code at 8 GPUs -- configs[2] (weak scaling, the headline), configs[3] (strong scaling of the fixed
# 6.4e6-spectra job), configs[4] (1e8 sets over 8 GPUs) with the timed all-gather of the band-mean table and one sharded inversion
# (tensor-core candidate pass).
mkdir -p gpurun_out; O=gpurun_out/r2n
run() {  # run <ngpus> <devices> <port> <outfile> <bench args...>
  local n=$1 dev=$2 port=$3 out=$4; shift 4
  CUDA_VISIBLE_DEVICES=$dev timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n "$@" > $out 2> $out.err
}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > ${O}_gpus.txt
run 8 0,1,2,3,4,5,6,7 29513 ${O}_bench_c2_n8.json --steps 10 --warmup 3 --no-cpu
run 8 0,1,2,3,4,5,6,7 29514 ${O}_bench_c4_n8.json --config 4 --steps 3 --warmup 3 --no-cpu
run 8 0,1,2,3,4,5,6,7 29515 ${O}_bench_c3_n8.json --config 3 --steps 5 --warmup 3 --no-cpu
for f in ${O}_bench_c2_n8 ${O}_bench_c4_n8 ${O}_bench_c3_n8; do echo "== $f"; tail -1 $f.json | cut -c1-3000; tail -3 $f.json.err | cut -c1-300; done
