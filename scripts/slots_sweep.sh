#!/bin/bash
# ring length of compress_ws_kernel (tuning builds with -DNDZB_WS_SLOTS=n) x debug mode (0 = full kernel, 7 = loads only)
for lib in tuning_s7 tuning_s9 tuning_s11 tuning; do
  for d in 0 7 3; do
    echo -n "$lib debug $d: "
    NDZB_LIB=build/exp/libndzb_$lib.so timeout 200 python scripts/kernel_time.py cfg2,cfg5 10 NDZB_WS_DEBUG=$d 2>&1 | grep -v "^ws stats" | sed -E 's/ \| decompress.*//' | tr '\n' ';'
    echo
  done
done
