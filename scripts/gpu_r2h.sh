# round 2, eighth GPU call: replicated class counters, micro timing with a non-empty queue, sanitizer passes
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2h_pytest_gpu.log); tail -4 gpurun_out/r2h_pytest_gpu.log | cut -c1-300
for dt in fp32 bf16; do (timeout 600 python bench.py --micro --micro-batch 64 --micro-dtype $dt > gpurun_out/r2h_micro_$dt.json 2> gpurun_out/r2h_micro_$dt.err); python -c "
import json; d=json.load(open('gpurun_out/r2h_micro_$dt.json')); k=d['config']['kernels']
for n,v in k.items():
    if 'ATen' not in n: print('   $dt %-82s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2h_micro_$dt.err; done
(timeout 600 python bench.py --micro --micro-batch 16 --micro-dtype fp32 > gpurun_out/r2h_micro16_fp32.json 2> gpurun_out/r2h_micro16_fp32.err); python -c "
import json; d=json.load(open('gpurun_out/r2h_micro16_fp32.json')); k=d['config']['kernels']
for n,v in k.items(): print('   B16 %-82s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2h_micro16_fp32.err
for ovf in default 0; do if [ $ovf = default ]; then unset ROBSEG_LOSS_GENERIC_OVF; else export ROBSEG_LOSS_GENERIC_OVF=$ovf; fi
(timeout 300 python bench.py --micro --micro-batch 24 --classes 21 --size 473 > gpurun_out/r2h_micro_voc473_ovf$ovf.json 2> gpurun_out/r2h_micro_voc473_ovf$ovf.err); python -c "
import json; d=json.load(open('gpurun_out/r2h_micro_voc473_ovf$ovf.json')); k=d['config']['kernels']
for n in ('loss_grad/mask-ce-avg','loss_grad/js-avg','loss_only/mask-ce-avg','argmax','apgd_step','pixel_hist/counts uniform-random','upsample_fwd x4 (ours)','upsample_bwd x4 (ours)'):
    v=k[n]; print('   voc473 ovf=$ovf %-34s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2h_micro_voc473_ovf$ovf.err; done
unset ROBSEG_LOSS_GENERIC_OVF
(timeout 300 python bench.py --micro --micro-batch 24 --classes 21 --size 472 > gpurun_out/r2h_micro_voc472.json 2>/dev/null); python -c "
import json; k=json.load(open('gpurun_out/r2h_micro_voc472.json'))['config']['kernels']
for n in ('loss_grad/mask-ce-avg','loss_only/mask-ce-avg','argmax'):
    v=k[n]; print('   voc472 (TMA) %-28s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))"
(timeout 300 python bench.py --micro --micro-batch 16 --classes 150 --size 473 > gpurun_out/r2h_micro_c150_473.json 2>/dev/null; ROBSEG_LOSS_GENERIC_OVF=0 timeout 300 python bench.py --micro --micro-batch 16 --classes 150 --size 473 > gpurun_out/r2h_micro_c150_473_ovf0.json 2>/dev/null); python -c "
import json
for f in ('gpurun_out/r2h_micro_c150_473.json','gpurun_out/r2h_micro_c150_473_ovf0.json'):
    k=json.load(open(f))['config']['kernels']; v=k['loss_grad/mask-ce-avg']; print('   c150 473^2', f[-14:], v['ms'], v['GBps'], v['frac'])"
(timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err); python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); c=d['config']; print(d['value'], d['ms_per_step'], d['e2e']['value'], c.get('fused_x4_variant',{}).get('value'), c.get('graph_variant',{}).get('value'), c['kernels_ms_per_step'], c['reference_on_gpu'].get('value'), d['roofline']['frac'], d['roofline']['avg_launch_ms'])" || tail -5 gpurun_out/r2h_bench.err
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r2h_sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize pass done' gpurun_out/r2h_sanitize_$tool.log | tr '\n' ' ')"
done
