mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -30 > gpurun_out/pytest_gpu.log)
tail -4 gpurun_out/pytest_gpu.log
L=gpurun_out/sweep4.log
: > $L
run() { echo "== $1 | $2" >> $L; env $1 timeout 60 python scripts/gpu_debug_hang.py $2 2>&1 | grep -E "run 2|rror" >> $L; }
run "A=1" "16 150 512 mask-ce-avg fp32"
run "A=1" "16 151 512 js-avg fp32"
run "A=1" "16 64 512 mask-ce-avg fp32"
run "ROBSEG_LOSS_VEC=2" "16 64 512 mask-ce-avg fp32"
run "A=1" "8 256 512 mask-ce-avg fp32"
run "ROBSEG_LOSS_VEC=2" "8 256 512 mask-ce-avg fp32"
run "A=1" "32 150 512 mask-ce-avg bf16"
run "ROBSEG_LOSS_VEC=2" "32 150 512 mask-ce-avg bf16"
run "ROBSEG_LOSS_VEC=8" "32 150 512 mask-ce-avg bf16"
cat $L
(timeout 1200 python bench.py --steps 2 --warmup 1 --debug-stack 240 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err)
head -c 4000 gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err
