mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log)
(timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -60 > gpurun_out/pytest_gpu.log)
(timeout 600 python bench.py --micro --micro-batch 16 > gpurun_out/micro16.json 2> gpurun_out/micro16.err)
(timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err)
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/micro16.json | head -c 3000; tail -3 gpurun_out/micro16.err; cat gpurun_out/bench_short.json | head -c 3000; tail -5 gpurun_out/bench_short.err
