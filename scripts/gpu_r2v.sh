# round 2, session 3, call 1: random-dispatch sweep of the loss entry points, full GPU suite, fresh ncu traffic captures,
# kernel-level brackets (robseg_profile_next_kernel) in the micro and the default bench
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fuzz.py -m gpu -q --timeout 200 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -60 > gpurun_out/r2v_fuzz.log); tail -3 gpurun_out/r2v_fuzz.log | cut -c1-300
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x --deselect tests/test_gpu_fuzz.py 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2v_pytest_gpu.log); tail -2 gpurun_out/r2v_pytest_gpu.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_tma_kernel -s 1 -c 1 -o gpurun_out/r2v_loss_c150 -f python scripts/loss_probe.py 16 150 512 mask-ce-avg fp32 > gpurun_out/r2v_ncu1.log 2>&1; tail -1 gpurun_out/r2v_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_tma_kernel -s 1 -c 1 -o gpurun_out/r2v_loss_c151 -f python scripts/loss_probe.py 16 151 512 mask-ce-avg fp32 > gpurun_out/r2v_ncu2.log 2>&1; tail -1 gpurun_out/r2v_ncu2.log
(timeout 300 python bench.py --micro --micro-batch 16 > gpurun_out/r2v_micro_b16.json 2> gpurun_out/r2v_micro_b16.err); python -c "
import json; k=json.load(open('gpurun_out/r2v_micro_b16.json'))['config']['kernels']
for n,v in k.items():
    if 'kernel_ms' in v: print('   %-60s call %8.4f ms %.3f | kernel %8.4f ms %.3f' % (n[:60], v['ms'], v['frac'], v['kernel_ms'], v['kernel_frac']))" || tail -5 gpurun_out/r2v_micro_b16.err
(timeout 900 python bench.py --steps 3 --warmup 3 --no-ref-on-gpu --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err); python -c "
import json; d=json.load(open('gpurun_out/r2v_bench.json')); c=d['config']; print(d['value'], d['e2e']['value'], d['gpu_launches'], c['kernels_ms_per_step'], c.get('loss_kernel_c151')); print(json.dumps(d['roofline']))" || tail -5 gpurun_out/r2v_bench.err
