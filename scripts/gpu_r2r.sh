# round 2, 17th GPU call (2 GPUs): the default bench under torchrun exactly as the driver launches it, both arms
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 600 $TR --master-port 29521 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2r_ref_n2.json 2> gpurun_out/r2r_ref_n2.err); tail -c 400 gpurun_out/r2r_ref_n2.json; echo
(timeout 900 $TR --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err); python -c "
import json; d=json.load(open('gpurun_out/r2r_bench_n2.json')); c=d['config']; print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['roofline']['frac'], d['clocks'], c['parallelism'])" || tail -20 gpurun_out/r2r_bench_n2.err
