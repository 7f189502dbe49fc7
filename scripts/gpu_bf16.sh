mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_tma -s 2 -c 1 -o gpurun_out/prof_loss_bf16 -f python scripts/loss_probe.py 32 150 512 mask-ce-avg bf16 > gpurun_out/ncu_bf16.log 2>&1; tail -2 gpurun_out/ncu_bf16.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_tma -s 2 -c 1 -o gpurun_out/prof_lossonly_bf16 -f python scripts/loss_probe_nograd.py 32 150 512 mask-ce-avg 0 bf16 > gpurun_out/ncu_bf16b.log 2>&1; tail -2 gpurun_out/ncu_bf16b.log
