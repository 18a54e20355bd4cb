#!/bin/bash
# Full GPU validation of a build: smoke, every -m gpu test group, CLI tests, stress, default bench, all BASELINE
# workloads, launch list + one ncu capture of the compress kernel. Writes gpurun_out/.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "[smoke] rc=$? $(tail -1 gpurun_out/f_smoke.log)"
bash scripts/gpu_tests.sh
timeout 600 python -m pytest tests/test_cli.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_cli.log 2>&1; echo "[cli] rc=$? $(tail -1 gpurun_out/pytest_cli.log)"
NDZB_WS_CHECK=1 timeout 90 python scripts/ws_stress.py float32 67108864 60 2>&1 | tail -1 | cut -c1-200
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "[bench] rc=$? $(python scripts/bench_summary.py gpurun_out/f_bench.json 2>/dev/null | head -3)"
for wl in cfg1 cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/f_bench_$wl.json 2> gpurun_out/f_bench_$wl.err
  echo "[bench $wl] rc=$? $(python scripts/bench_summary.py gpurun_out/f_bench_$wl.json 2>/dev/null | head -3)"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:compress|decompress|border|fixup|offset" -c 40 --csv --log-file gpurun_out/r3_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/f_ncu_list.log 2>&1; echo "[ncu list] rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:compress_ws -s 4 -c 1 -o gpurun_out/r3_compress_ws -f \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/f_ncu_c.log 2>&1; echo "[ncu compress] rc=$?"
