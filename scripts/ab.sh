#!/bin/bash
# A/B timing of several builds of the library on ONE box: scripts/ab.sh "cfg2,cfg3" reps lib1.so lib2.so ...
# ("default" = ndzip_b200/libndzip_b200.so). Two rounds, interleaved, so that drift shows.
wl=$1; reps=$2; shift 2
for round in 1 2; do
  for lib in "$@"; do
    if [ "$lib" = default ]; then NDZB_LIB= timeout 300 python scripts/kernel_time.py $wl $reps 2>&1 | grep -v "^ws stats"
    else NDZB_LIB=$lib timeout 300 python scripts/kernel_time.py $wl $reps 2>&1 | grep -v "^ws stats"; fi
  done
done
