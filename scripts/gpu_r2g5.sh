# round 2, last call: smoke() and a one-step default bench on the final tree (graph variants included), the trainer test once more
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g5_smoke.log 2>&1); tail -1 gpurun_out/r2g5_smoke.log
(timeout 300 python -m pytest tests/test_gpu_dropin.py -m gpu -q --timeout 300 -k run_train_main 2>&1 | tail -1)
(timeout 600 python bench.py --steps 1 --warmup 3 --no-ref-on-gpu --no-cpu-baseline > gpurun_out/r2g5_bench.json 2> gpurun_out/r2g5_bench.err); python -c "
import json; d=json.load(open('gpurun_out/r2g5_bench.json')); c=d['config']; print(d['value'], d['e2e']['value'], d['roofline']['frac'], [c.get(k,{}).get('value') for k in ('fused_x4_variant','graph_variant','graph_fused_x4_variant')])" || tail -5 gpurun_out/r2g5_bench.err
