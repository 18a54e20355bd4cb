#!/bin/bash
timeout 300 python scripts/ws_time.py cfg2 20 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=4 2>&1 | grep -E "avg|Error"
for v in 1 2 4; do NDZB_WS_STATS=1 NDZB_WS_VARIANT=$v timeout 120 python scripts/ws_time.py cfg2 5 2>&1 | grep "ws stats" | tail -2; done
