mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?"
tail -c 1500 gpurun_out/bench_n2.err; head -c 600 gpurun_out/bench_n2.json
