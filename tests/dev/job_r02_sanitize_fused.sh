mkdir -p gpurun_out/r02
for tool in memcheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 python tests/dev/sanitize_small.py > gpurun_out/r02/san_fused_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY' gpurun_out/r02/san_fused_$tool.log | tail -1)"
done
