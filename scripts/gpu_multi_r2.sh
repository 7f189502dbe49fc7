# round 2 multi-GPU evidence: configs[2] (Segmenter, n_iter 300, image-sharded), configs[3] (PIR-AT under DDP) and
# configs[4] (kernel microbench on every GPU at once).  usage: bash scripts/gpu_multi_r2.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r2m_gpus_n$N.txt
(NCCL_DEBUG=INFO timeout 900 $TR --master-port 29511 bench.py --gpus $N --workload pirat --steps 5 --warmup 3 > gpurun_out/r2m_pirat_n$N.json 2> gpurun_out/r2m_pirat_n$N.err); grep -E "Init COMPLETE|NVLS|nranks" gpurun_out/r2m_pirat_n$N.err | head -12 > gpurun_out/r2m_nccl_n$N.txt; python -c "
import json; d=json.load(open('gpurun_out/r2m_pirat_n$N.json')); c=d['config']; print('pirat N=$N', d['value'], d['ms_per_step'], c['modes_images_per_s'], c['attack_time_allreduce_bytes_per_step'])" || tail -5 gpurun_out/r2m_pirat_n$N.err
(timeout 1200 $TR --master-port 29512 bench.py --gpus $N --model segmenter --n-iter 300 --batch 4 --steps 1 --warmup 3 > gpurun_out/r2m_segmenter300_n$N.json 2> gpurun_out/r2m_segmenter300_n$N.err); python -c "
import json; d=json.load(open('gpurun_out/r2m_segmenter300_n$N.json')); c=d['config']; print('segmenter n_iter=300 N=$N', d['value'], d['ms_per_step'], d['e2e']['value'], c['kernels_ms_per_step'])" || tail -5 gpurun_out/r2m_segmenter300_n$N.err
for dt in fp32 bf16; do (timeout 600 $TR --master-port 29513 bench.py --gpus $N --micro --micro-batch 64 --micro-dtype $dt > gpurun_out/r2m_micro_${dt}_n$N.json 2> gpurun_out/r2m_micro_${dt}_n$N.err); python -c "
import json; d=json.load(open('gpurun_out/r2m_micro_${dt}_n$N.json')); k=d['config']['kernels']
print('micro $dt N=$N aggregate', d['value'], 'GB/s')
for n in ('loss_grad/mask-ce-avg','loss_only/mask-ce-avg','argmax','apgd_step','pixel_hist/counts uniform-random','pixel_hist/full uniform-random'):
    v=k[n]; print('   %-36s %8.4f ms %8.1f GB/s/GPU (slowest %s, aggregate %s)' % (n, v['ms'], v['GBps'], v.get('GBps_slowest_rank'), v.get('GBps_aggregate')))" || tail -5 gpurun_out/r2m_micro_${dt}_n$N.err; done
