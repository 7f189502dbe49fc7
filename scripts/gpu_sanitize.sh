mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize pass done' gpurun_out/sanitize_$tool.log | tr '\n' ' ')"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_generic -s 2 -c 1 -o gpurun_out/prof_loss_generic python scripts/loss_probe_nograd.py 24 21 473 mask-ce-avg 1 > gpurun_out/ncu_generic.log 2>&1; tail -1 gpurun_out/ncu_generic.log
