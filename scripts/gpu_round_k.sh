#!/bin/bash
for v in 6 7; do NDZB_WS_STATS=1 NDZB_WS_VARIANT=$v timeout 120 python scripts/ws_time.py cfg2 5 2>&1 | grep "ws stats" | tail -2; done
NDZB_WS_STATS=1 NDZB_WS_VARIANT=4 timeout 120 python scripts/ws_time.py cfg3 5 2>&1 | grep "ws stats" | tail -2
