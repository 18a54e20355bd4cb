#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider --tb=short \
    -k "all_kernels or misaligned_stream or aligned_multi_cube" > gpurun_out/j_pytest_ws.log 2>&1; echo "[ws tests] rc=$? $(tail -1 gpurun_out/j_pytest_ws.log)"
echo "--- timing (cfg2)"
timeout 300 python scripts/ws_time.py cfg2 20 NDZB_COMPRESS_KERNEL=v1 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=3 NDZB_WS_VARIANT=4 NDZB_WS_VARIANT=5 2>&1 | grep -E "avg|Error"
for v in 6 7; do NDZB_WS_STATS=1 NDZB_WS_VARIANT=$v timeout 120 python scripts/ws_time.py cfg2 5 2>&1 | grep "ws stats" | tail -1; done
echo "--- timing (cfg3)"
timeout 300 python scripts/ws_time.py cfg3 20 NDZB_COMPRESS_KERNEL=v1 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=3 2>&1 | grep -E "avg|Error"
NDZB_WS_STATS=1 NDZB_WS_VARIANT=4 timeout 120 python scripts/ws_time.py cfg3 5 2>&1 | grep "ws stats" | tail -1
echo "--- timing (cfg5, cfg1)"
timeout 300 python scripts/ws_time.py cfg5 10 NDZB_COMPRESS_KERNEL=v1 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=4 2>&1 | grep -E "avg|Error"
timeout 300 python scripts/ws_time.py cfg1 20 NDZB_COMPRESS_KERNEL=v1 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=4 2>&1 | grep -E "avg|Error"
