#!/bin/bash
# Full validation of the round's state: all -m gpu tests, smoke, default bench, launch list + ncu captures.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; echo "[smoke] rc=$? $(tail -1 gpurun_out/l_smoke.log)"
bash scripts/gpu_tests.sh
timeout 600 python bench.py > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "[bench] rc=$? $(python scripts/bench_summary.py gpurun_out/l_bench.json 2>/dev/null | head -5)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/l_ncu_list.log 2>&1; echo "[ncu list] rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:compress_ws -s 4 -c 1 -o gpurun_out/r2_compress_ws -f \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/l_ncu_c.log 2>&1; echo "[ncu compress] rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decompress_kernel -s 4 -c 1 -o gpurun_out/r2_decompress -f \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/l_ncu_d.log 2>&1; echo "[ncu decompress] rc=$?"
