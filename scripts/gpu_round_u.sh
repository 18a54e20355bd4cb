#!/bin/bash
# load pipeline alone (NDZB_WS_DEBUG=4: no encoding; 6 = also no look-back; 7 = also no copy)
for wl in cfg2 cfg5 cfg3; do
  timeout 300 python scripts/ws_time.py $wl 10 NDZB_WS_VARIANT=0 NDZB_WS_VARIANT=0,NDZB_WS_DEBUG=4 NDZB_WS_VARIANT=0,NDZB_WS_DEBUG=6 NDZB_WS_VARIANT=0,NDZB_WS_DEBUG=7 2>&1 | grep -E "avg|Error"
done
NDZB_WS_STATS=1 NDZB_WS_VARIANT=4 NDZB_WS_DEBUG=6 timeout 120 python scripts/ws_time.py cfg2 5 2>&1 | grep "ws stats" | tail -2
