# round 2, 18th GPU call: register-resident small-C loss variant (C <= 24)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2s_pytest_gpu.log); tail -4 gpurun_out/r2s_pytest_gpu.log | cut -c1-300
for creg in 1 0; do for S in 473 472; do
(ROBSEG_LOSS_CREG=$creg timeout 300 python bench.py --micro --micro-batch 24 --classes 21 --size $S > gpurun_out/r2s_micro_voc${S}_creg$creg.json 2> gpurun_out/r2s_micro_voc${S}_creg$creg.err); python -c "
import json; d=json.load(open('gpurun_out/r2s_micro_voc${S}_creg$creg.json')); k=d['config']['kernels']
for n,v in k.items():
    if ('loss' in n or 'argmax' in n) and 'ATen' not in n: print('   voc$S creg=$creg %-82s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2s_micro_voc${S}_creg$creg.err; done; done
(timeout 300 python bench.py --batch 2 --classes 21 --eps 4 --steps 5 --warmup 3 --no-cpu-baseline --no-ref-on-gpu > gpurun_out/r2s_b2c21.json 2> gpurun_out/r2s_b2c21.err); python -c "
import json; d=json.load(open('gpurun_out/r2s_b2c21.json')); c=d['config']; print('B2 C21', d['value'], d['ms_per_step'], c['kernels_ms_per_step'], d['roofline']['frac'])" || tail -5 gpurun_out/r2s_b2c21.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_generic_ovf -s 1 -c 1 -o gpurun_out/r2s_loss_ovf_creg -f python scripts/loss_probe.py 24 21 473 mask-ce-avg fp32 > gpurun_out/r2s_ncu1.log 2>&1; tail -1 gpurun_out/r2s_ncu1.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r2s_sanitize_memcheck.log 2>&1; echo "== memcheck: $(grep -E 'ERROR SUMMARY|sanitize pass done' gpurun_out/r2s_sanitize_memcheck.log | tr '\n' ' ')"
