# round 2, N=1 verification + the N=1 points of the multi-GPU workloads
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 400 2>&1 | grep -v "^E   +\|where <built-in\|where tensor" | tail -40 > gpurun_out/r2d_pytest_gpu.log); tail -8 gpurun_out/r2d_pytest_gpu.log | cut -c1-300
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1); tail -2 gpurun_out/r2d_smoke.log
for a in "16 150 128 4" "16 150 32 16" "16 150 64 8" "2 21 128 4" "16 150 128 4 js-avg" "16 151 128 4"; do timeout 200 python scripts/loss_up_probe.py $a 2>&1 | tail -1 | tee -a gpurun_out/r2d_loss_up_probe.log; done
bash scripts/gpu_multi_r2.sh 1
