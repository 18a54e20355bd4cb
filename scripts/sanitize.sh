#!/bin/bash
# compute-sanitizer over small runs of the default kernels (compress_ws_kernel + decompress_kernel via smoke(), and a
# 600-cube 1-D run): memcheck, synccheck, racecheck. Writes gpurun_out/sanitize_*.log.
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_${tool}_smoke.log 2>&1
  echo "[$tool smoke] rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/sanitize_${tool}_smoke.log | tr '\n' ' ')"
done
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python scripts/ws_stress.py float32 2457600 2 > gpurun_out/sanitize_racecheck_1d.log 2>&1
echo "[racecheck 1d 600 cubes] rc=$? $(grep -E 'RACECHECK SUMMARY|^ok' gpurun_out/sanitize_racecheck_1d.log | tr '\n' ' ')"
