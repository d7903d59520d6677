mkdir -p gpurun_out/r02
for tool in memcheck racecheck initcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tests/dev/sanitize_lookahead.py > gpurun_out/r02/san_la_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02/san_la_$tool.log | tail -1)"
done
