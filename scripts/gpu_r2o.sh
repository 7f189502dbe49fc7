# round 2, 14th GPU call: final verification after restoring the on-demand walk forward -- tests, smoke, VOC micro, default bench
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2o_pytest_gpu.log); tail -4 gpurun_out/r2o_pytest_gpu.log | cut -c1-300
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1); tail -2 gpurun_out/r2o_smoke.log
(timeout 300 python bench.py --micro --micro-batch 24 --classes 21 --size 473 > gpurun_out/r2o_micro_voc473.json 2> gpurun_out/r2o_micro_voc473.err); python -c "
import json; d=json.load(open('gpurun_out/r2o_micro_voc473.json')); k=d['config']['kernels']
for n,v in k.items():
    if 'upsample' in n: print('   voc473 %-82s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2o_micro_voc473.err
(timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err); python -c "
import json; d=json.load(open('gpurun_out/r2o_bench.json')); c=d['config']; print(d['value'], d['ms_per_step'], d['e2e']['value'], c.get('fused_x4_variant',{}).get('value'), c.get('graph_variant',{}).get('value'), c['kernels_ms_per_step'], c['reference_on_gpu'].get('value'), d['roofline']['frac'], d['roofline']['avg_launch_ms'], d.get('cpu_baseline',{}).get('value'), c.get('loss_kernel_c151'))" || tail -5 gpurun_out/r2o_bench.err
(timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2o_bench_ref.json 2> gpurun_out/r2o_bench_ref.err); tail -c 600 gpurun_out/r2o_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv --log-file gpurun_out/r2o_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-on-gpu > gpurun_out/r2o_bench_under_ncu.json 2> gpurun_out/r2o_bench_under_ncu.err
wc -l gpurun_out/r2o_launches_bench.csv
