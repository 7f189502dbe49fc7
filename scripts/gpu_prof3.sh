# r01b pass: GPU tests, default bench line, ncu captures of the histogram / up-sampling kernels,
# launch list of the bench, micro-benchmarks
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -5 > gpurun_out/pytest_gpu.log); tail -2 gpurun_out/pytest_gpu.log
(timeout 900 python bench.py > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err); tail -c 600 gpurun_out/bench_r01b.json
(timeout 600 python bench.py --micro --micro-batch 64 > gpurun_out/micro64_fp32.json 2> gpurun_out/micro64_fp32.err)
(timeout 600 python bench.py --micro --micro-batch 16 > gpurun_out/micro16_fp32.json 2> gpurun_out/micro16_fp32.err)
python - <<'PY'
import json
for n in ("micro64_fp32","micro16_fp32"):
    d=json.load(open(f'gpurun_out/{n}.json'))
    for k,v in d['config']['kernels'].items(): print(f"{n} {k:45s} {v['ms']:8.4f} ms {v['GBps']:8.1f} GB/s {v['frac']:.3f}")
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pixel_|upsample_" -c 14 -o gpurun_out/prof_hist_up -f python bench.py --micro --micro-batch 64 --steps 1 --warmup 1 > gpurun_out/ncu_hist_up.log 2>&1
tail -2 gpurun_out/ncu_hist_up.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv --log-file gpurun_out/launches_bench_r01b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu3.json 2> gpurun_out/bench_under_ncu3.err
wc -l gpurun_out/launches_bench_r01b.csv
