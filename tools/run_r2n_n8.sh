mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2n_bench_c2_n8.json 2> gpurun_out/r2n_bench_c2_n8.json.err
tail -1 gpurun_out/r2n_bench_c2_n8.json | cut -c1-300
