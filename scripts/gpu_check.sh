mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -8 > gpurun_out/pytest_gpu.log); tail -3 gpurun_out/pytest_gpu.log
(timeout 600 python bench.py --micro --micro-batch 64 > gpurun_out/micro64_fp32.json 2> gpurun_out/micro64_fp32.err)
python - <<'PY'
import json
d=json.load(open('gpurun_out/micro64_fp32.json'))
for k,v in d['config']['kernels'].items(): print(f"{k:45s} {v['ms']:8.4f} ms {v['GBps']:8.1f} GB/s {v['frac']:.3f}")
PY
