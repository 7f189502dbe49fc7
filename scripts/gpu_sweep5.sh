mkdir -p gpurun_out
L=gpurun_out/sweep5.log
: > $L
run() { echo "== $1 | $2" >> $L; env $1 timeout 60 python scripts/gpu_debug4.py $2 2>&1 | tail -1 >> $L; }
run "A=1" "24 21 473 mask-ce-avg 1"
run "A=1" "24 21 472 mask-ce-avg 1"
run "A=1" "16 151 473 mask-ce-avg 1"
run "A=1" "16 150 512 mask-ce-avg 0"
run "ROBSEG_LOSS_VEC=1" "16 150 512 mask-ce-avg 0"
run "ROBSEG_LOSS_VEC=1 ROBSEG_LOSS_SLOTS=1" "16 150 512 mask-ce-avg 0"
run "ROBSEG_LOSS_G=2" "16 150 512 mask-ce-avg 0"
run "ROBSEG_LOSS_G=2 ROBSEG_LOSS_VEC=1" "16 150 512 mask-ce-avg 0"
cat $L
