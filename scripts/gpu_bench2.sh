mkdir -p gpurun_out
(timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err)
(timeout 1200 python bench.py --steps 3 --warmup 3 --stock-upsample --no-cpu-baseline > gpurun_out/bench_stock.json 2> gpurun_out/bench_stock.err)
python - <<'PY'
import json
for n in ("bench_fast","bench_stock"):
    try:
        d=json.load(open(f"gpurun_out/{n}.json"))
        print(n, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "attack_ms", d["config"]["attack_side_ms_per_step"], d["config"]["kernels_ms_per_step"], d["roofline"]["achieved"], d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(n, "ERR", e)
PY
tail -3 gpurun_out/bench_fast.err
