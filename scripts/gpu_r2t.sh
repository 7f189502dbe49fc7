# round 2, final GPU call: the driver's sequence on the final tree -- full GPU suite, smoke, reference arm, our arm
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2t_pytest_gpu.log); tail -4 gpurun_out/r2t_pytest_gpu.log | cut -c1-300
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1); tail -2 gpurun_out/r2t_smoke.log
(timeout 300 python bench.py --micro --micro-batch 24 --classes 21 --size 472 > gpurun_out/r2t_micro_voc472.json 2>/dev/null); python -c "
import json; k=json.load(open('gpurun_out/r2t_micro_voc472.json'))['config']['kernels']
for n in ('loss_grad/mask-ce-avg','loss_grad/js-avg','loss_only/mask-ce-avg','argmax'):
    v=k[n]; print('   voc472 %-28s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))"
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2t_bench_ref.json 2> gpurun_out/r2t_bench_ref.err); python -c "
import json; d=json.load(open('gpurun_out/r2t_bench_ref.json')); print('reference arm', d['value'], d['ms_per_step'], d['cpu_baseline']['cores'])"
(timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err); python -c "
import json; d=json.load(open('gpurun_out/r2t_bench.json')); c=d['config']; print(d['value'], d['ms_per_step'], d['e2e']['value'], c.get('fused_x4_variant',{}).get('value'), c.get('graph_variant',{}).get('value'), c['kernels_ms_per_step'], c['reference_on_gpu'].get('value'), d['roofline']['frac'], d['roofline']['avg_launch_ms'], d.get('cpu_baseline',{}).get('value'), c.get('loss_kernel_c151',{}).get('frac'), d['clocks'])" || tail -5 gpurun_out/r2t_bench.err
