#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider --tb=short \
    -k "stream_equals_oracle or aligned_multi or golden_single" 2>&1 | tail -1
for wl in cfg2 cfg3 cfg5; do
  timeout 200 python scripts/ws_time.py $wl 20 NDZB_WS_VARIANT=0 2>&1 | grep -E "avg|Error"
  timeout 200 python scripts/dec_time.py $wl 20 2>&1 | grep -E "avg|Error"
done
