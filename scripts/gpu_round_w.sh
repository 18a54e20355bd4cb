#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider --tb=short \
    -k "all_kernels and float32" 2>&1 | tail -1
timeout 300 python scripts/ws_time.py cfg2 10 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=3 2>&1 | grep -E "avg|Error"
NDZB_WS_STATS=1 NDZB_WS_VARIANT=4 timeout 120 python scripts/ws_time.py cfg2 5 2>&1 | grep "ws stats" | tail -2
timeout 300 python scripts/ws_time.py cfg5 10 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 2>&1 | grep -E "avg|Error"
