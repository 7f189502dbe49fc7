mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout 200 -k "loss_kernel or custom_ops or cross_entropy or pgd_attack or apgd" 2>&1 | tail -12 > gpurun_out/pytest_generic.log); tail -6 gpurun_out/pytest_generic.log
for v in 4 2 1; do echo "ROBSEG_LOSS_GENERIC_VEC=$v"; ROBSEG_LOSS_GENERIC_VEC=$v python bench.py --micro --micro-batch 24 --classes 21 --size 473 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
for k,v in d['config']['kernels'].items():
    if k.startswith('loss') or k=='argmax': print(f'  {k:30s} {v[\"ms\"]:8.4f} ms {v[\"GBps\"]:8.1f} GB/s {v[\"frac\"]:.3f}')"; done
