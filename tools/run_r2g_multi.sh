#!/bin/bash
# round 2g (8-GPU box): configs[3] strong scaling at 2/4/8 GPUs, configs[4] (1e8 sets over 8 GPUs) with the timed all-gather of
# the band-mean table and one sharded inversion, and the host<->device copy ceiling with 1..8 GPUs concurrently.
mkdir -p gpurun_out; O=gpurun_out/r2g; 
run() {  # run <ngpus> <devices> <port> <outfile> <bench args...>
  local n=$1 dev=$2 port=$3 out=$4; shift 4
  CUDA_VISIBLE_DEVICES=$dev timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n "$@" > $out 2> $out.err
}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > ${O}_gpus.txt
# phase A: N=2 and N=4 side by side on disjoint GPUs (no e2e: the host cores are shared)
run 2 0,1 29511 ${O}_c3_n2.json --config 3 --steps 5 --warmup 3 --no-cpu --no-e2e &
run 4 2,3,4,5 29512 ${O}_c3_n4.json --config 3 --steps 5 --warmup 3 --no-cpu --no-e2e &
wait
# phase B: all 8 GPUs
run 8 0,1,2,3,4,5,6,7 29513 ${O}_c3_n8.json --config 3 --steps 5 --warmup 3 --no-cpu
run 8 0,1,2,3,4,5,6,7 29514 ${O}_c4_n8.json --config 4 --steps 3 --warmup 3 --no-cpu
tools/micro/bin/d2h_probe 8 1024 6 > ${O}_d2h_probe_8gpu.txt 2>&1
for f in ${O}_c3_n2 ${O}_c3_n4 ${O}_c3_n8 ${O}_c4_n8; do echo "== $f"; tail -1 $f.json | cut -c1-2500; tail -3 $f.json.err | cut -c1-300; done
cat ${O}_d2h_probe_8gpu.txt
