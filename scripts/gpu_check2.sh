mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -8 > gpurun_out/pytest_gpu.log); tail -3 gpurun_out/pytest_gpu.log
for a in "24 21 473 mask-ce-avg 1" "16 151 473 mask-ce-avg 1" "24 21 472 mask-ce-avg 1"; do timeout 60 python scripts/gpu_debug4.py $a 2>&1 | tail -1; done
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2)
