# round 2, fifth GPU call: new x2 forward, bf16 keep-e loss kernel, counters fused into the loss kernels,
# bench with graph variant + e2e warm-up fix
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2e_pytest_gpu.log); tail -8 gpurun_out/r2e_pytest_gpu.log | cut -c1-300
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1); tail -2 gpurun_out/r2e_smoke.log
(timeout 300 python scripts/upsample_probe.py 2>&1 | grep -v Warn > gpurun_out/r2e_upsample_probe.log); cat gpurun_out/r2e_upsample_probe.log
(ROBSEG_UP_X2_CELL=1 timeout 300 python scripts/upsample_probe.py 2>&1 | grep -v Warn > gpurun_out/r2e_upsample_probe_old.log); grep "64,64\|32,32\]->64\|16,16\]->32" gpurun_out/r2e_upsample_probe_old.log
for dt in fp32 bf16; do (timeout 600 python bench.py --micro --micro-batch 64 --micro-dtype $dt > gpurun_out/r2e_micro_$dt.json 2> gpurun_out/r2e_micro_$dt.err); python -c "
import json; d=json.load(open('gpurun_out/r2e_micro_$dt.json')); k=d['config']['kernels']
for n,v in k.items():
    if 'ATen' not in n: print('   $dt %-58s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2e_micro_$dt.err; done
(timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err); python -c "
import json; d=json.load(open('gpurun_out/r2e_bench.json')); c=d['config']; print(d['value'], d['ms_per_step'], d['e2e'], c.get('fused_x4_variant'), c.get('graph_variant'), c['kernels_ms_per_step'], c['reference_on_gpu'].get('value'), d['roofline'])" || tail -5 gpurun_out/r2e_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:upsample_fwd_x2 -s 2 -c 1 -o gpurun_out/r2e_up_x2 -f python scripts/upsample_probe.py > gpurun_out/r2e_ncu1.log 2>&1; tail -1 gpurun_out/r2e_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_tma_kernel -s 1 -c 1 -o gpurun_out/r2e_loss_bf16 -f python scripts/loss_probe.py 64 150 512 mask-ce-avg bf16 > gpurun_out/r2e_ncu2.log 2>&1; tail -1 gpurun_out/r2e_ncu2.log
