mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -15 > gpurun_out/pytest_gpu.log)
tail -4 gpurun_out/pytest_gpu.log
(timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err)
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r01.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "attack_ms", d["config"]["attack_side_ms_per_step"], d["config"]["attack_side_frac_of_step"], d["config"]["kernels_ms_per_step"], d["roofline"]["achieved"], d.get("cpu_baseline",{}), d["clocks"], d["gpu_launches"])
PY
tail -3 gpurun_out/bench_r01.err
