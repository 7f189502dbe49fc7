mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -5 > gpurun_out/pytest_gpu.log); tail -2 gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"apgd_step|upsample_.*x4|pixel_hist" -s 12 -c 8 -o gpurun_out/prof_others python bench.py --micro --micro-batch 16 --steps 5 --warmup 3 > gpurun_out/ncu_others.log 2>&1
tail -2 gpurun_out/ncu_others.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu2.json 2> gpurun_out/bench_under_ncu2.err
wc -l gpurun_out/launches_bench_final.csv
(timeout 600 python bench.py --micro --micro-batch 64 > gpurun_out/micro64_fp32.json 2> gpurun_out/micro64_fp32.err)
(timeout 600 python bench.py --micro --micro-batch 64 --micro-dtype bf16 > gpurun_out/micro64_bf16.json 2> gpurun_out/micro64_bf16.err)
python - <<'PY'
import json
for n in ("micro64_fp32","micro64_bf16"):
    d=json.load(open(f'gpurun_out/{n}.json'))
    for k,v in d['config']['kernels'].items(): print(f"{n} {k:40s} {v['ms']:8.4f} ms {v['GBps']:8.1f} GB/s {v['frac']:.3f}")
PY
