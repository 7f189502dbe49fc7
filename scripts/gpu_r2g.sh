# round 2, seventh GPU call: deferred class counting, half trips, faster over-fetch fill; VOC-shape micro per path;
# default bench (roofline with the fused counters in the step)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2g_pytest_gpu.log); tail -6 gpurun_out/r2g_pytest_gpu.log | cut -c1-300
for dt in fp32 bf16; do (timeout 600 python bench.py --micro --micro-batch 64 --micro-dtype $dt > gpurun_out/r2g_micro_$dt.json 2> gpurun_out/r2g_micro_$dt.err); python -c "
import json; d=json.load(open('gpurun_out/r2g_micro_$dt.json')); k=d['config']['kernels']
for n,v in k.items():
    if 'ATen' not in n and 'pixel_hist' not in n and 'upsample' not in n: print('   $dt %-82s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2g_micro_$dt.err; done
for ovf in default 4 1 0; do if [ $ovf = default ]; then unset ROBSEG_LOSS_GENERIC_OVF; else export ROBSEG_LOSS_GENERIC_OVF=$ovf; fi
(timeout 300 python bench.py --micro --micro-batch 24 --classes 21 --size 473 > gpurun_out/r2g_micro_voc473_ovf$ovf.json 2> gpurun_out/r2g_micro_voc473_ovf$ovf.err); python -c "
import json; d=json.load(open('gpurun_out/r2g_micro_voc473_ovf$ovf.json')); k=d['config']['kernels']
for n in ('loss_grad/mask-ce-avg','loss_grad/js-avg','loss_only/mask-ce-avg','argmax','upsample_fwd x4 (ours)','upsample_bwd x4 (ours)'):
    v=k[n]; print('   voc473 ovf=$ovf %-28s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2g_micro_voc473_ovf$ovf.err; done
unset ROBSEG_LOSS_GENERIC_OVF
(timeout 300 python bench.py --micro --micro-batch 24 --classes 21 --size 472 > gpurun_out/r2g_micro_voc472.json 2>/dev/null); python -c "
import json; k=json.load(open('gpurun_out/r2g_micro_voc472.json'))['config']['kernels']
for n in ('loss_grad/mask-ce-avg','loss_only/mask-ce-avg','argmax'):
    v=k[n]; print('   voc472 (TMA) %-28s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))"
(timeout 300 python bench.py --micro --micro-batch 16 --classes 150 --size 473 > gpurun_out/r2g_micro_c150_473.json 2>/dev/null; ROBSEG_LOSS_GENERIC_OVF=0 timeout 300 python bench.py --micro --micro-batch 16 --classes 150 --size 473 > gpurun_out/r2g_micro_c150_473_ovf0.json 2>/dev/null); python -c "
import json
for f in ('gpurun_out/r2g_micro_c150_473.json','gpurun_out/r2g_micro_c150_473_ovf0.json'):
    k=json.load(open(f))['config']['kernels']; v=k['loss_grad/mask-ce-avg']; print('   c150 473^2', f[-14:], v['ms'], v['GBps'], v['frac'])"
(timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err); python -c "
import json; d=json.load(open('gpurun_out/r2g_bench.json')); c=d['config']; print(d['value'], d['ms_per_step'], d['e2e']['value'], c.get('fused_x4_variant',{}).get('value'), c.get('graph_variant',{}).get('value'), c['kernels_ms_per_step'], c['reference_on_gpu'].get('value'), d['roofline']['frac'], d['roofline']['avg_launch_ms'])" || tail -5 gpurun_out/r2g_bench.err
(timeout 900 python bench.py --steps 2 --warmup 3 --pred-maps --no-ref-on-gpu --no-cpu-baseline > gpurun_out/r2g_bench_predmaps.json 2> gpurun_out/r2g_bench_predmaps.err); python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_predmaps.json')); c=d['config']; print('pred-maps', d['value'], d['ms_per_step'], c['kernels_ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])" || tail -5 gpurun_out/r2g_bench_predmaps.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_generic_ovf -s 1 -c 1 -o gpurun_out/r2g_loss_ovf -f python scripts/loss_probe.py 24 21 473 mask-ce-avg fp32 > gpurun_out/r2g_ncu1.log 2>&1; tail -1 gpurun_out/r2g_ncu1.log
