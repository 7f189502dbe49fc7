mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/ddp_pirat_smoke.py > gpurun_out/ddp_pirat.log 2>&1); tail -3 gpurun_out/ddp_pirat.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err)
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'])"
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | head -c 600)
