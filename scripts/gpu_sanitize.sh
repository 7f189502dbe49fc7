mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize pass done' gpurun_out/sanitize_$tool.log | tr '\n' ' ')"
done
