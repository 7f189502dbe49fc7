mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/sea_sharded_check.py > gpurun_out/sea_sharded.log 2>&1; echo "sharded rc=$?"; grep -E "sharded SEA|Error|error|assert" gpurun_out/sea_sharded.log | head -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2>/dev/null; echo "ref rc=$?"; head -c 400 gpurun_out/bench_ref_n2.json
