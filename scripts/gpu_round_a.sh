#!/bin/bash
# First GPU pass for the warp-specialised compress kernel: smoke, new parity tests, A/B sweep, one ncu capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "[smoke] rc=$? $(tail -1 gpurun_out/a_smoke.log)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider --tb=short \
    -k "all_kernels or misaligned_stream or aligned_multi_cube" > gpurun_out/a_pytest_ws.log 2>&1; echo "[ws tests] rc=$? $(tail -1 gpurun_out/a_pytest_ws.log)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider --tb=short \
    -k "baseline_config or sharded or offloader or context" > gpurun_out/a_pytest_base.log 2>&1; echo "[baseline tests] rc=$? $(tail -1 gpurun_out/a_pytest_base.log)"
WORKLOADS="cfg2" VARIANTS="0 1 2 3 4" scripts/ws_sweep.sh
cp gpurun_out/ws_sweep.jsonl gpurun_out/ws_sweep_cfg2.jsonl
WORKLOADS="cfg3 cfg1" VARIANTS="0 1 2 3" scripts/ws_sweep.sh
cp gpurun_out/ws_sweep.jsonl gpurun_out/ws_sweep_cfg31.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:compress_ws -s 4 -c 1 -o gpurun_out/r2a_compress_ws -f \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/a_ncu.log 2>&1; echo "[ncu] rc=$?"
