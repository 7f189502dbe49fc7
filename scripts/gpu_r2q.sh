# round 2, 16th GPU call: SegMenter drop-in test, cheaper counting path; counts probes; segmenter bench
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2q_pytest_gpu.log); tail -4 gpurun_out/r2q_pytest_gpu.log | cut -c1-300
python scripts/counts_probe.py 24 21 473 coherent 2>&1 | tail -2 | tee gpurun_out/r2q_counts_probe.log
python scripts/counts_probe.py 24 21 473 2>&1 | tail -2 | tee -a gpurun_out/r2q_counts_probe.log
python scripts/counts_probe.py 16 150 512 2>&1 | tail -2 | tee -a gpurun_out/r2q_counts_probe.log
(timeout 600 python bench.py --model segmenter --n-iter 30 --batch 4 --steps 2 --warmup 3 --no-ref-on-gpu > gpurun_out/r2q_segmenter.json 2> gpurun_out/r2q_segmenter.err); python -c "
import json; d=json.load(open('gpurun_out/r2q_segmenter.json')); c=d['config']; print(d['value'], d['ms_per_step'], c['kernels_ms_per_step'], d['roofline'])" || tail -5 gpurun_out/r2q_segmenter.err
(timeout 600 python bench.py --model segmenter --n-iter 30 --batch 4 --steps 2 --warmup 3 --no-ref-on-gpu --pred-maps > gpurun_out/r2q_segmenter_predmaps.json 2> gpurun_out/r2q_segmenter_predmaps.err); python -c "
import json; d=json.load(open('gpurun_out/r2q_segmenter_predmaps.json')); c=d['config']; print('pred-maps', d['value'], d['ms_per_step'], c['kernels_ms_per_step'])" || tail -5 gpurun_out/r2q_segmenter_predmaps.err
