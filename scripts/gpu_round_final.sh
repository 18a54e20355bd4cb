#!/bin/bash
# End-of-session validation: smoke, every -m gpu test, default bench, all BASELINE workloads, sanitizer on smoke.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "[smoke] rc=$? $(tail -1 gpurun_out/f_smoke.log)"
bash scripts/gpu_tests.sh
timeout 600 python -m pytest tests/test_cli.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_cli.log 2>&1; echo "[cli] rc=$? $(tail -1 gpurun_out/pytest_cli.log)"
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "[bench] rc=$? $(python scripts/bench_summary.py gpurun_out/f_bench.json 2>/dev/null | head -3)"
for wl in cfg1 cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/f_bench_$wl.json 2> gpurun_out/f_bench_$wl.err
  echo "[bench $wl] rc=$? $(python scripts/bench_summary.py gpurun_out/f_bench_$wl.json 2>/dev/null | head -3)"
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_sanitizer.log 2>&1; echo "[memcheck smoke] rc=$? $(grep -E 'ERROR SUMMARY|smoke ok' gpurun_out/f_sanitizer.log | tr '\n' ' ')"
