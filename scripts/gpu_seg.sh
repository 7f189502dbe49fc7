mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 150 -k "segmenter or upsample" 2>&1 | tail -8 > gpurun_out/pytest_seg.log); tail -4 gpurun_out/pytest_seg.log
(timeout 600 python bench.py --model segmenter --no-cpu-baseline > gpurun_out/bench_segmenter.json 2> gpurun_out/bench_segmenter.err); tail -c 1500 gpurun_out/bench_segmenter.json; tail -3 gpurun_out/bench_segmenter.err
(timeout 600 python bench.py --model segmenter --no-cpu-baseline --stock-upsample > gpurun_out/bench_segmenter_stock.json 2> gpurun_out/bench_segmenter_stock.err); head -c 400 gpurun_out/bench_segmenter_stock.json
