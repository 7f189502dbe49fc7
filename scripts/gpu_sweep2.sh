mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -30 > gpurun_out/pytest_gpu.log)
tail -5 gpurun_out/pytest_gpu.log
L=gpurun_out/sweep2.log
: > $L
run() { echo "== $1 | $2" >> $L; env $1 timeout 60 python scripts/gpu_debug_hang.py $2 2>&1 | grep -E "run 2|rror" >> $L; }
for e in "ROBSEG_LOSS_G=1" "ROBSEG_LOSS_G=2" "ROBSEG_LOSS_G=2 ROBSEG_LOSS_WARPS=8" "ROBSEG_LOSS_G=2 ROBSEG_LOSS_WARPS=6 ROBSEG_LOSS_SLOTS=2" "ROBSEG_LOSS_G=1 ROBSEG_LOSS_WARPS=6 ROBSEG_LOSS_SLOTS=2"; do
  run "$e" "16 150 512 mask-ce-avg fp32"
done
run "ROBSEG_LOSS_G=2" "16 150 512 js-avg fp32"
run "ROBSEG_LOSS_G=2" "64 150 512 mask-ce-avg fp32"
for e in "ROBSEG_LOSS_G=1" "ROBSEG_LOSS_G=1 ROBSEG_LOSS_SLOTS=1" "ROBSEG_LOSS_G=1 ROBSEG_LOSS_SLOTS=2"; do
  run "$e" "64 21 512 mask-ce-avg fp32"
done
for e in "ROBSEG_LOSS_G=1" "ROBSEG_LOSS_G=2"; do
  run "$e" "16 64 512 mask-ce-avg fp32"
  run "$e" "32 150 512 mask-ce-avg bf16"
done
cat $L
