#!/bin/bash
timeout 300 python scripts/ws_time.py cfg4 10 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=3 2>&1 | grep -E "avg|Error"
timeout 300 python scripts/ws_time.py cfg3 10 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=1 NDZB_WS_VARIANT=2 NDZB_WS_VARIANT=3 2>&1 | grep -E "avg|Error"
NDZB_WS_STATS=1 NDZB_WS_VARIANT=4 timeout 120 python scripts/ws_time.py cfg4 5 2>&1 | grep "ws stats" | tail -2
