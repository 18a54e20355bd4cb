#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider --tb=short \
    -k "stream_equals_oracle or baseline_config or aligned_multi or golden_single" > gpurun_out/m_pytest.log 2>&1; echo "[decode tests] rc=$? $(tail -1 gpurun_out/m_pytest.log)"
for wl in cfg2 cfg1 cfg3 cfg5; do timeout 200 python scripts/dec_time.py $wl 20 NDZB_DEC_CTAS=5 NDZB_DEC_CTAS=6 2>&1 | grep -E "avg|Error"; done
