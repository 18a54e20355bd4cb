#!/bin/bash
# Short GPU check after a kernel change: ws parity tests, sharded / offloader / context tests, stress, timings.
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 --timeout-method=thread -p no:cacheprovider --tb=short \
    -k "all_kernels or aligned_multi or misaligned_stream or sharded or offloader or context or repeated_launches" 2>&1 | tail -1
NDZB_WS_CHECK=1 timeout 90 python scripts/ws_stress.py float32 67108864 40 2>&1 | tail -1 | cut -c1-200
for wl in cfg2 cfg5 cfg1; do timeout 200 python scripts/ws_time.py $wl 20 NDZB_WS_VARIANT=0 2>&1 | grep -E "avg|Error"; done
