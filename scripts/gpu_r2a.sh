# round 2, first GPU call: the real drop-in tests + the full -m gpu suite + the three bench workloads
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q -s --timeout 400 2>&1 | tail -40 > gpurun_out/r2a_pytest_dropin.log); tail -25 gpurun_out/r2a_pytest_dropin.log
(timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_gpu_dropin.py 2>&1 | tail -12 > gpurun_out/r2a_pytest_gpu.log); tail -4 gpurun_out/r2a_pytest_gpu.log
(timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err); tail -c 3000 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
(timeout 600 python bench.py --workload pirat --steps 3 --warmup 3 > gpurun_out/r2a_pirat.json 2> gpurun_out/r2a_pirat.err); tail -c 2500 gpurun_out/r2a_pirat.json; tail -3 gpurun_out/r2a_pirat.err
(timeout 600 python bench.py --model segmenter --steps 1 --warmup 3 > gpurun_out/r2a_segmenter.json 2> gpurun_out/r2a_segmenter.err); tail -c 2500 gpurun_out/r2a_segmenter.json; tail -3 gpurun_out/r2a_segmenter.err
