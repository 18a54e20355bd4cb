#!/bin/bash
mkdir -p gpurun_out
echo "--- timing (cfg2): retire-warp sweep"
timeout 300 python scripts/ws_time.py cfg2 20 NDZB_COMPRESS_KERNEL=v1 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=3 NDZB_WS_VARIANT=4 NDZB_WS_VARIANT=5 NDZB_WS_VARIANT=6 NDZB_WS_VARIANT=7 2>&1 | grep -E "avg|Error"
NDZB_WS_STATS=1 NDZB_WS_VARIANT=7 timeout 120 python scripts/ws_time.py cfg2 5 2>&1 | grep "ws stats" | tail -1
echo "--- timing (cfg3)"
timeout 300 python scripts/ws_time.py cfg3 20 NDZB_COMPRESS_KERNEL=v1 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=3 NDZB_WS_VARIANT=4 2>&1 | grep -E "avg|Error"
NDZB_WS_STATS=1 NDZB_WS_VARIANT=4 timeout 120 python scripts/ws_time.py cfg3 5 2>&1 | grep "ws stats" | tail -1
echo "--- timing (cfg5, cfg1)"
timeout 300 python scripts/ws_time.py cfg5 10 NDZB_COMPRESS_KERNEL=v1 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=3 NDZB_WS_VARIANT=5 2>&1 | grep -E "avg|Error"
timeout 300 python scripts/ws_time.py cfg1 20 NDZB_COMPRESS_KERNEL=v1 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=3 NDZB_WS_VARIANT=5 2>&1 | grep -E "avg|Error"
