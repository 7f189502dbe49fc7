# round 2, final GPU call of the last session: the driver's sequence on the final tree (full GPU suite, smoke, reference arm,
# our arm), memcheck over the kernels changed in this session, the micro table, the ncu launch list of the bench command
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2f_pytest_gpu.log); tail -2 gpurun_out/r2f_pytest_gpu.log | cut -c1-300
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1); tail -1 gpurun_out/r2f_smoke.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r2f_sanitize_memcheck.log 2>&1
echo "== memcheck: $(grep -E 'ERROR SUMMARY|sanitize pass done' gpurun_out/r2f_sanitize_memcheck.log | tr '\n' ' ')"
(timeout 400 python bench.py --micro > gpurun_out/r2f_micro.json 2> gpurun_out/r2f_micro.err); python -c "
import json; k=json.load(open('gpurun_out/r2f_micro.json'))['config']['kernels']
for n,v in k.items(): print('   %-70s %8.4f ms %8.1f GB/s %.3f %s' % (n[:70], v['ms'], v['GBps'], v['frac'], ('kernel %.4f ms %.3f' % (v['kernel_ms'], v['kernel_frac'])) if 'kernel_ms' in v else ''))" || tail -5 gpurun_out/r2f_micro.err
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err); python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_ref.json')); print('reference arm', d['value'], d['ms_per_step'], d['cpu_baseline']['cores'])"
(timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err); python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json')); c=d['config']; print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], [c.get(k,{}).get('value') for k in ('fused_x4_variant','graph_variant','graph_fused_x4_variant')], c['kernels_ms_per_step'], c['reference_on_gpu'].get('value'), d.get('cpu_baseline',{}).get('value'), c.get('loss_kernel_c151'), d['clocks']); print(json.dumps(d['roofline']))" || tail -5 gpurun_out/r2f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 1 --warmup 1 --no-ref-on-gpu --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1; wc -l gpurun_out/r2f_launches.csv
