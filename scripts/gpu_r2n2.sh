# round 2, last session: the default bench under torch.distributed.run on 2 GPUs (final tree)
mkdir -p gpurun_out
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2n2_bench.json 2> gpurun_out/r2n2_bench.err); python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2n2_bench.json') if l.startswith('{')][-1]); c=d['config']; print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['launches_timed'], d['clocks'], c['parallelism'])" || tail -20 gpurun_out/r2n2_bench.err
