# round 2, 12th GPU call: interval-walk forward, counter replicas zeroed by a kernel instead of cudaMemsetAsync
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2l_pytest_gpu.log); tail -4 gpurun_out/r2l_pytest_gpu.log | cut -c1-300
for z in kernel memset; do if [ $z = memset ]; then export ROBSEG_COUNTS_MEMSET=1; else unset ROBSEG_COUNTS_MEMSET; fi
(timeout 300 python bench.py --micro --micro-batch 24 --classes 21 --size 473 > gpurun_out/r2l_micro_voc473_$z.json 2> gpurun_out/r2l_micro_voc473_$z.err); python -c "
import json; d=json.load(open('gpurun_out/r2l_micro_voc473_$z.json')); k=d['config']['kernels']
for n,v in k.items():
    if 'ATen' not in n and 'pixel_hist' not in n: print('   voc473 zero=$z %-82s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2l_micro_voc473_$z.err; done
unset ROBSEG_COUNTS_MEMSET
(timeout 600 python bench.py --micro --micro-batch 16 --micro-dtype fp32 > gpurun_out/r2l_micro16_fp32.json 2> gpurun_out/r2l_micro16_fp32.err); python -c "
import json; d=json.load(open('gpurun_out/r2l_micro16_fp32.json')); k=d['config']['kernels']
for n,v in k.items():
    if 'loss' in n or 'argmax' in n: print('   B16 %-82s %8.4f ms %8.1f GB/s %.3f' % (n, v['ms'], v['GBps'], v['frac']))" || tail -5 gpurun_out/r2l_micro16_fp32.err
python - <<'PY' 2>&1 | grep -v Warn | tee gpurun_out/r2l_upsample_probe_voc.log
import importlib, statistics, sys, torch
sys.path.insert(0, ".")
ops = importlib.import_module("robust-segmentation_b200.ops")
dev = torch.device("cuda:0"); g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, inner):
    for _ in range(2): fn()
    ts = []
    for _ in range(5):
        flush.zero_(); torch.cuda._sleep(1_000_000)
        ops.profile_start(); fn(); torch.cuda.synchronize()
        ts.append(sum(ms for n, _, ms in ops.profile_stop() if n == inner))
    return statistics.median(ts)
for B, C, s, S in [(24, 21, 119, 473), (24, 512, 59, 119), (24, 512, 29, 119), (24, 512, 14, 119), (24, 512, 29, 59), (24, 512, 14, 29)]:
    low = torch.randn(B, C, s, s, device=dev, generator=g); gup = torch.randn(B, C, S, S, device=dev, generator=g)
    nb = 4 * (low.numel() + gup.numel())
    f, b = t(lambda: ops._upsample_fwd(low, S, S), "upsample_fwd"), t(lambda: ops._upsample_bwd(gup, s, s), "upsample_bwd")
    print(f"[{B},{C},{s},{s}]->{S}: fwd {f*1e3:7.1f} us {nb/f/1e6:6.0f} GB/s  bwd {b*1e3:7.1f} us {nb/b/1e6:6.0f} GB/s", flush=True)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:upsample_fwd_walk -s 2 -c 1 -o gpurun_out/r2l_up_fwd_walk -f python bench.py --micro --micro-batch 24 --classes 21 --size 473 > gpurun_out/r2l_ncu1.log 2>&1; tail -1 gpurun_out/r2l_ncu1.log
