#!/bin/bash
mkdir -p gpurun_out
echo "--- timing (cfg2): v1, production ws, stats twins with prefetch limits"
timeout 300 python scripts/ws_time.py cfg2 20 NDZB_COMPRESS_KERNEL=v1 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=3 NDZB_WS_VARIANT=4 NDZB_WS_VARIANT=5 NDZB_WS_VARIANT=6 NDZB_WS_VARIANT=7 2>&1 | grep -E "avg|Error"
echo "--- stats (cfg2)"
for v in 1 2 3 5 6 7; do
  echo "variant $v"; NDZB_WS_STATS=1 NDZB_WS_VARIANT=$v timeout 120 python scripts/ws_time.py cfg2 5 2>&1 | grep "ws stats" | tail -1
done
echo "variant 1, no look-back"; NDZB_WS_STATS=1 NDZB_WS_VARIANT=1 NDZB_WS_DEBUG=2 timeout 120 python scripts/ws_time.py cfg2 5 2>&1 | grep "ws stats" | tail -1
echo "--- timing (cfg3)"
timeout 300 python scripts/ws_time.py cfg3 20 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=3 NDZB_WS_VARIANT=4 2>&1 | grep -E "avg|Error"
for v in 1 2; do
  echo "variant $v"; NDZB_WS_STATS=1 NDZB_WS_VARIANT=$v timeout 120 python scripts/ws_time.py cfg3 5 2>&1 | grep "ws stats" | tail -1
done
