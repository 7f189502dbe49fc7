# round 2, first GPU call: the real drop-in tests + the full -m gpu suite + the three bench workloads
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q -s --timeout 400 2>&1 | tail -40 > gpurun_out/r2a_pytest_dropin.log); tail -25 gpurun_out/r2a_pytest_dropin.log
(timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_gpu_dropin.py 2>&1 | tail -12 > gpurun_out/r2a_pytest_gpu.log); tail -4 gpurun_out/r2a_pytest_gpu.log
(timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err); tail -c 3000 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
(timeout 600 python bench.py --workload pirat --steps 3 --warmup 3 > gpurun_out/r2a_pirat.json 2> gpurun_out/r2a_pirat.err); tail -c 2500 gpurun_out/r2a_pirat.json; tail -3 gpurun_out/r2a_pirat.err
(timeout 600 python bench.py --model segmenter --steps 1 --warmup 3 > gpurun_out/r2a_segmenter.json 2> gpurun_out/r2a_segmenter.err); tail -c 2500 gpurun_out/r2a_segmenter.json; tail -3 gpurun_out/r2a_segmenter.err
(timeout 600 python -m pytest tests/test_gpu_fused_upsample.py -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/r2a_pytest_fused.log); tail -15 gpurun_out/r2a_pytest_fused.log
for a in "16 150 128 4" "16 150 32 16" "16 150 64 8" "2 21 128 4" "16 150 128 4 js-avg"; do timeout 200 python scripts/loss_up_probe.py $a 2>&1 | tail -1 | tee -a gpurun_out/r2a_loss_up_probe.log; done
