# round 2, last session: the other two bench entry points on the final tree (short runs)
mkdir -p gpurun_out
(timeout 400 python bench.py --model segmenter --n-iter 30 --batch 4 --steps 2 --warmup 3 --no-ref-on-gpu > gpurun_out/r2g4_segmenter.json 2> gpurun_out/r2g4_segmenter.err); python -c "
import json; d=json.load(open('gpurun_out/r2g4_segmenter.json')); c=d['config']; print('segmenter', d['value'], d['e2e']['value'], d['gpu_launches'], c['kernels_ms_per_step']); print(json.dumps(d['roofline'])[:700])" || tail -5 gpurun_out/r2g4_segmenter.err
(timeout 400 python bench.py --workload pirat --steps 2 --warmup 3 > gpurun_out/r2g4_pirat.json 2> gpurun_out/r2g4_pirat.err); python -c "
import json; d=json.load(open('gpurun_out/r2g4_pirat.json')); c=d['config']; print('pirat', d['value'], d['e2e']['value'], d['gpu_launches'], {k: v for k, v in c.items() if 'images' in k or 'ms' in k}); print(json.dumps(d['roofline'])[:500])" || tail -5 gpurun_out/r2g4_pirat.err
