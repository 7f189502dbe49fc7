mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -30 > gpurun_out/pytest_gpu.log)
tail -5 gpurun_out/pytest_gpu.log
L=gpurun_out/sweep_loss.log
: > $L
for cfg in "1 12" "2 6" "3 4" "1 8" "1 6"; do
  set -- $cfg
  echo "== C=150 slots=$1 warps=$2" >> $L
  ROBSEG_LOSS_SLOTS=$1 ROBSEG_LOSS_WARPS=$2 timeout 60 python scripts/gpu_debug_hang.py 16 150 512 mask-ce-avg fp32 2>&1 | grep -E "run 2|Error|error" >> $L
done
for cfg in "1 15" "2 10" "3 7" "4 5"; do
  set -- $cfg
  echo "== C=21 slots=$1 warps=$2" >> $L
  ROBSEG_LOSS_SLOTS=$1 ROBSEG_LOSS_WARPS=$2 timeout 60 python scripts/gpu_debug_hang.py 64 21 512 mask-ce-avg fp32 2>&1 | grep -E "run 2|Error|error" >> $L
done
for args in "4 150 512 argmax 1 0 0" "4 150 512 argmax 0 1 0" "16 150 512 js-avg 0 1 1" "16 150 512 mask-ce-avg 0 1 0"; do
  echo "== $args" >> $L
  timeout 60 python scripts/gpu_debug3.py $args 2>&1 | tail -3 >> $L
done
cat $L
timeout 400 ncu --set full --clock-control none --import-source on -k regex:loss_tma -s 2 -c 1 -o gpurun_out/prof_loss_c150 python scripts/gpu_debug_hang.py 4 150 512 mask-ce-avg fp32 > gpurun_out/ncu_c150.log 2>&1
tail -3 gpurun_out/ncu_c150.log
ls -la gpurun_out
