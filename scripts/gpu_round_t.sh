#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider --tb=short \
    -k "all_kernels or aligned_multi or misaligned_stream" > gpurun_out/t_pytest_ws.log 2>&1; echo "[ws tests] rc=$? $(tail -1 gpurun_out/t_pytest_ws.log)"
NDZB_WS_CHECK=1 NDZB_WS_VARIANT=1 timeout 90 python scripts/ws_stress.py float32 67108864 40 2>&1 | tail -1 | cut -c1-200
timeout 300 python scripts/ws_time.py cfg2 20 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=3 2>&1 | grep -E "avg|Error"
NDZB_WS_STATS=1 NDZB_WS_VARIANT=4 timeout 120 python scripts/ws_time.py cfg2 5 2>&1 | grep "ws stats" | tail -2
timeout 300 python scripts/ws_time.py cfg3 20 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 2>&1 | grep -E "avg|Error"
timeout 300 python scripts/ws_time.py cfg5 10 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 2>&1 | grep -E "avg|Error"
timeout 300 python scripts/ws_time.py cfg1 20 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 2>&1 | grep -E "avg|Error"
