mkdir -p gpurun_out
(timeout 900 python bench.py --steps 3 --warmup 3 --no-ref-on-gpu --no-cpu-baseline > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err); python -c "
import json; d=json.load(open('gpurun_out/r2u_bench.json')); c=d['config']; print(d['value'], d['e2e']['value'], c['kernels_ms_per_step']); print(json.dumps(d['roofline']))" || tail -5 gpurun_out/r2u_bench.err
(timeout 600 python bench.py --model segmenter --n-iter 30 --batch 4 --steps 2 --warmup 3 --no-ref-on-gpu > gpurun_out/r2u_segmenter.json 2> gpurun_out/r2u_segmenter.err); python -c "
import json; d=json.load(open('gpurun_out/r2u_segmenter.json')); c=d['config']; print(d['value'], c['kernels_ms_per_step']); print(json.dumps(d['roofline']))" || tail -5 gpurun_out/r2u_segmenter.err
(timeout 900 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | tail -2)
