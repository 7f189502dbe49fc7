mkdir -p gpurun_out
(timeout 900 python bench.py --micro --micro-batch 64 > gpurun_out/micro64_fp32.json 2> gpurun_out/micro64_fp32.err)
(timeout 900 python bench.py --micro --micro-batch 64 --micro-dtype bf16 > gpurun_out/micro64_bf16.json 2> gpurun_out/micro64_bf16.err)
(timeout 600 python bench.py --micro --micro-batch 16 > gpurun_out/micro16_fp32.json 2> gpurun_out/micro16_fp32.err)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:loss_tma -s 2 -c 1 -o gpurun_out/prof_loss_c150_default python scripts/gpu_debug_hang.py 16 150 512 mask-ce-avg fp32 > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 3000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 16 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
head -c 1500 gpurun_out/micro64_fp32.json; echo; head -c 1500 gpurun_out/micro64_bf16.json; echo; tail -2 gpurun_out/micro64_fp32.err; wc -l gpurun_out/launches_bench.csv
