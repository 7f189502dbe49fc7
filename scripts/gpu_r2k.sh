# round 2, 11th GPU call: class counting deferred by one tile in the generic and fused-up-sampling kernels too
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2k_pytest_gpu.log); tail -4 gpurun_out/r2k_pytest_gpu.log | cut -c1-300
(timeout 300 python bench.py --micro --micro-batch 24 --classes 21 --size 473 > gpurun_out/r2k_micro_voc473.json 2> gpurun_out/r2k_micro_voc473.err); python -c "
import json; d=json.load(open('gpurun_out/r2k_micro_voc473.json')); k=d['config']['kernels']
for n,v in k.items():
    if 'ATen' not in n and 'pixel_hist' not in n: print('   voc473 %-82s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2k_micro_voc473.err
(timeout 300 python bench.py --micro --micro-batch 24 --classes 21 --size 472 > gpurun_out/r2k_micro_voc472.json 2> /dev/null); python -c "
import json; d=json.load(open('gpurun_out/r2k_micro_voc472.json')); k=d['config']['kernels']
for n,v in k.items():
    if 'loss' in n or 'argmax' in n: print('   voc472 %-82s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))"
for dt in fp32 bf16; do (timeout 600 python bench.py --micro --micro-batch 64 --micro-dtype $dt > gpurun_out/r2k_micro_$dt.json 2> gpurun_out/r2k_micro_$dt.err); python -c "
import json; d=json.load(open('gpurun_out/r2k_micro_$dt.json')); k=d['config']['kernels']
for n,v in k.items():
    if 'loss' in n or 'argmax' in n: print('   $dt %-82s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2k_micro_$dt.err; done
for a in "16 150 32 16" "16 150 128 4" "4 150 32 16"; do timeout 200 python scripts/loss_up_probe.py $a 2>&1 | tail -1 | tee -a gpurun_out/r2k_loss_up_probe.log; done
(timeout 600 python bench.py --model segmenter --n-iter 30 --batch 4 --steps 2 --warmup 3 --no-ref-on-gpu > gpurun_out/r2k_segmenter.json 2> gpurun_out/r2k_segmenter.err); python -c "
import json; d=json.load(open('gpurun_out/r2k_segmenter.json')); c=d['config']; print(d['value'], d['ms_per_step'], c['kernels_ms_per_step'], d['roofline'])" || tail -5 gpurun_out/r2k_segmenter.err
