# round 2, 15th GPU call: sanitizer over the final kernels + fresh ncu --set full captures of the final loss kernels
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r2p_sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize pass done' gpurun_out/r2p_sanitize_$tool.log | tr '\n' ' ')"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_tma_kernel -s 1 -c 1 -o gpurun_out/r2p_loss_c150 -f python scripts/loss_probe.py 16 150 512 mask-ce-avg fp32 > gpurun_out/r2p_ncu1.log 2>&1; tail -1 gpurun_out/r2p_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_tma_kernel -s 1 -c 1 -o gpurun_out/r2p_loss_c151 -f python scripts/loss_probe.py 16 151 512 mask-ce-avg fp32 > gpurun_out/r2p_ncu2.log 2>&1; tail -1 gpurun_out/r2p_ncu2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_tma_kernel -s 1 -c 1 -o gpurun_out/r2p_loss_bf16 -f python scripts/loss_probe.py 64 150 512 mask-ce-avg bf16 > gpurun_out/r2p_ncu3.log 2>&1; tail -1 gpurun_out/r2p_ncu3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_tma_kernel -s 3 -c 1 -o gpurun_out/r2p_loss_c150_counts -f python scripts/counts_probe.py 16 150 512 > gpurun_out/r2p_ncu4.log 2>&1; tail -1 gpurun_out/r2p_ncu4.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_up_kernel -s 3 -c 1 -o gpurun_out/r2p_loss_up_x16 -f python scripts/loss_up_probe.py 16 150 32 16 > gpurun_out/r2p_ncu5.log 2>&1; tail -1 gpurun_out/r2p_ncu5.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_up_kernel -s 3 -c 1 -o gpurun_out/r2p_loss_up_x4 -f python scripts/loss_up_probe.py 16 150 128 4 > gpurun_out/r2p_ncu6.log 2>&1; tail -1 gpurun_out/r2p_ncu6.log
