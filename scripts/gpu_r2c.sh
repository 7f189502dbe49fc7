# round 2, third GPU call: graph-iteration tests, config-1-shape graph vs eager, default bench (+ fused x4 variant),
# segmenter with the fused loss, ncu captures (loss_up x4/x16, loss_tma C=150/151), launch list of the default bench
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_graph_iteration.py "tests/test_gpu_parity.py::test_graphed_model_attack_equals_eager" tests/test_gpu_fused_upsample.py -m gpu -q --timeout 300 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -60 > gpurun_out/r2c_pytest_graph.log); tail -40 gpurun_out/r2c_pytest_graph.log | cut -c1-300
for v in "" "--graph"; do (timeout 600 python bench.py --batch 2 --classes 21 --eps 4 --steps 5 --warmup 3 --no-cpu-baseline --no-ref-on-gpu $v > gpurun_out/r2c_b2c21$v.json 2> gpurun_out/r2c_b2c21$v.err); python -c "
import json,sys; d=json.load(open('gpurun_out/r2c_b2c21$v.json')); print('B2 C21 $v', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])" || tail -5 gpurun_out/r2c_b2c21$v.err; done
(timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err); python -c "
import json; d=json.load(open('gpurun_out/r2c_bench.json')); c=d['config']; print(d['value'], d['ms_per_step'], c.get('fused_x4_variant'), c.get('peak_mem_GiB'), c['reference_on_gpu'].get('value'))" || tail -5 gpurun_out/r2c_bench.err
(timeout 600 python bench.py --model segmenter --steps 2 --warmup 3 > gpurun_out/r2c_segmenter.json 2> gpurun_out/r2c_segmenter.err); python -c "
import json; d=json.load(open('gpurun_out/r2c_segmenter.json')); c=d['config']; print(d['value'], d['ms_per_step'], c['kernels_ms_per_step'], c.get('peak_mem_GiB'), c['reference_on_gpu'].get('value'))" || tail -5 gpurun_out/r2c_segmenter.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_up_kernel -s 3 -c 1 -o gpurun_out/r2c_loss_up_x4 -f python scripts/loss_up_probe.py 16 150 128 4 > gpurun_out/r2c_ncu1.log 2>&1; tail -1 gpurun_out/r2c_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_up_kernel -s 3 -c 1 -o gpurun_out/r2c_loss_up_x16 -f python scripts/loss_up_probe.py 16 150 32 16 > gpurun_out/r2c_ncu2.log 2>&1; tail -1 gpurun_out/r2c_ncu2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_tma_kernel -s 1 -c 1 -o gpurun_out/r2c_loss_c151 -f python scripts/loss_probe.py 16 151 512 mask-ce-avg fp32 > gpurun_out/r2c_ncu3.log 2>&1; tail -1 gpurun_out/r2c_ncu3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_tma_kernel -s 1 -c 1 -o gpurun_out/r2c_loss_c150 -f python scripts/loss_probe.py 16 150 512 mask-ce-avg fp32 > gpurun_out/r2c_ncu4.log 2>&1; tail -1 gpurun_out/r2c_ncu4.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv --log-file gpurun_out/r2c_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-on-gpu > gpurun_out/r2c_bench_under_ncu.json 2> gpurun_out/r2c_bench_under_ncu.err
wc -l gpurun_out/r2c_launches_bench.csv
