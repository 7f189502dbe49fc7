# round 2, session 3, call 3: trainer APGD-branch call, graph + fused x4 variant of the default bench, configs[0] shape
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "trainer_apgd_branch or pirat_training" 2>&1 | tail -3)
(timeout 900 python bench.py --steps 2 --warmup 3 --no-ref-on-gpu --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err); python -c "
import json; d=json.load(open('gpurun_out/r2x_bench.json')); c=d['config']
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['traffic_capture'])
for k in ('fused_x4_variant','graph_variant','graph_fused_x4_variant'): print(k, {a:b for a,b in c[k].items() if a in ('value','ms_per_step','peak_mem_GiB','host_launch_calls','unavailable')})" || tail -5 gpurun_out/r2x_bench.err
(timeout 900 python bench.py --batch 2 --classes 21 --eps 4 --steps 3 --warmup 3 --no-ref-on-gpu --no-cpu-baseline > gpurun_out/r2x_b2c21.json 2> gpurun_out/r2x_b2c21.err); python -c "
import json; d=json.load(open('gpurun_out/r2x_b2c21.json')); c=d['config']
print('B2 C21', d['value'], d['e2e']['value'], d['roofline']['frac'], c['kernels_ms_per_step'])
for k in ('fused_x4_variant','graph_variant','graph_fused_x4_variant'): print(k, {a:b for a,b in c[k].items() if a in ('value','ms_per_step','peak_mem_GiB','host_launch_calls','unavailable')})" || tail -5 gpurun_out/r2x_b2c21.err
