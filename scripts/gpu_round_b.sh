#!/bin/bash
mkdir -p gpurun_out
echo "--- debug=1 (plain stores)"; NDZB_WS_DEBUG=1 timeout 120 python scripts/ws_repro.py 2>&1 | tail -3
echo "--- tma stores"; timeout 120 python scripts/ws_repro.py 2>&1 | tail -3
echo "--- sanitizer, debug=1"; NDZB_WS_DEBUG=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/ws_repro.py > gpurun_out/b_san_dbg.log 2>&1; grep -vE "^=+$" gpurun_out/b_san_dbg.log | grep -E "=========|words" | head -30
echo "--- sanitizer, tma stores"; timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/ws_repro.py > gpurun_out/b_san.log 2>&1; grep -E "=========|words" gpurun_out/b_san.log | head -40
