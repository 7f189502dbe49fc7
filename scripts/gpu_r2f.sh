# round 2, sixth GPU call: x2 forward (7 blocks/SM) + new x2 backward, leader-extraction counters, label prefetch,
# 16-byte over-fetch generic loss path at the VOC shape, bf16 schedule sweep
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2f_pytest_gpu.log); tail -8 gpurun_out/r2f_pytest_gpu.log | cut -c1-300
(timeout 300 python scripts/upsample_probe.py 2>&1 | grep -v Warn > gpurun_out/r2f_upsample_probe.log); cat gpurun_out/r2f_upsample_probe.log
for dt in fp32 bf16; do (timeout 600 python bench.py --micro --micro-batch 64 --micro-dtype $dt > gpurun_out/r2f_micro_$dt.json 2> gpurun_out/r2f_micro_$dt.err); python -c "
import json; d=json.load(open('gpurun_out/r2f_micro_$dt.json')); k=d['config']['kernels']
for n,v in k.items():
    if 'ATen' not in n and 'pixel_hist' not in n: print('   $dt %-58s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2f_micro_$dt.err; done
for ovf in default 4 2 1 0; do if [ $ovf = default ]; then unset ROBSEG_LOSS_GENERIC_OVF; else export ROBSEG_LOSS_GENERIC_OVF=$ovf; fi
(timeout 300 python bench.py --micro --micro-batch 24 --classes 21 --size 473 > gpurun_out/r2f_micro_voc473_ovf$ovf.json 2> gpurun_out/r2f_micro_voc473_ovf$ovf.err); python -c "
import json; d=json.load(open('gpurun_out/r2f_micro_voc473_ovf$ovf.json')); k=d['config']['kernels']
for n in ('loss_grad/mask-ce-avg','loss_grad/js-avg','loss_only/mask-ce-avg','argmax'):
    v=k[n]; print('   voc473 ovf=$ovf %-28s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2f_micro_voc473_ovf$ovf.err; done
unset ROBSEG_LOSS_GENERIC_OVF
(timeout 300 python bench.py --micro --micro-batch 16 --classes 150 --size 473 > gpurun_out/r2f_micro_c150_473.json 2>/dev/null; ROBSEG_LOSS_GENERIC_OVF=0 timeout 300 python bench.py --micro --micro-batch 16 --classes 150 --size 473 > gpurun_out/r2f_micro_c150_473_ovf0.json 2>/dev/null); python -c "
import json
for f in ('gpurun_out/r2f_micro_c150_473.json','gpurun_out/r2f_micro_c150_473_ovf0.json'):
    k=json.load(open(f))['config']['kernels']; v=k['loss_grad/mask-ce-avg']; print('   c150 473^2', f[-10:], v['ms'], v['GBps'], v['frac'])"
echo "bf16 schedule sweep (loss_probe 64 150 512 mask-ce-avg bf16: VEC WARPS SLOTS)"
for cfg in "2 11 1" "2 5 2" "2 8 1" "4 5 1" "4 3 1" "4 2 2"; do set -- $cfg; (ROBSEG_LOSS_VEC=$1 ROBSEG_LOSS_WARPS=$2 ROBSEG_LOSS_SLOTS=$3 timeout 120 python scripts/loss_probe.py 64 150 512 mask-ce-avg bf16 2>&1 | grep "run 2" | sed "s/^/   vec=$1 warps=$2 slots=$3 /") | tee -a gpurun_out/r2f_bf16_sweep.log; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_generic_ovf -s 1 -c 1 -o gpurun_out/r2f_loss_ovf -f python scripts/loss_probe.py 24 21 473 mask-ce-avg fp32 > gpurun_out/r2f_ncu1.log 2>&1; tail -1 gpurun_out/r2f_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:upsample_bwd_x2 -s 2 -c 1 -o gpurun_out/r2f_up_bwd_x2 -f python scripts/upsample_probe.py > gpurun_out/r2f_ncu2.log 2>&1; tail -1 gpurun_out/r2f_ncu2.log
