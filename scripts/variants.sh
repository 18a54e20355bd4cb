#!/bin/bash
# Times every tuning variant of compress_ws_kernel (a -DNDZB_TUNING build): scripts/variants.sh lib.so "cfg2,cfg3" reps "0 1 2 3"
lib=$1; wl=$2; reps=$3; shift 3
for v in $1; do
  echo -n "variant $v: "
  NDZB_LIB=$lib timeout 300 python scripts/kernel_time.py $wl $reps NDZB_WS_VARIANT=$v 2>&1 | grep -v "^ws stats" | sed -E 's/ \| decompress.*//' | tr '\n' ';'
  echo
done
