mkdir -p gpurun_out
(timeout 900 python bench.py --micro --micro-batch 16 > gpurun_out/micro16_fp32.json 2> gpurun_out/micro16_fp32.err)
python - <<'PY'
import json
d=json.load(open('gpurun_out/micro16_fp32.json'))
for k,v in d['config']['kernels'].items(): print(f"{k:45s} {v['ms']:9.4f} ms {v['GBps']:8.1f} GB/s {v['frac']:.3f}")
PY
tail -3 gpurun_out/micro16_fp32.err
