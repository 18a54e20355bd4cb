#!/bin/bash
# A/B of the compress kernels on one B200: v1 (compress_kernel) against the warp-specialised
# compress_ws_kernel in its (encoder groups, retire warps) variants. Writes gpurun_out/ws_sweep.jsonl.
mkdir -p gpurun_out
out=gpurun_out/ws_sweep.jsonl
: > $out
run() {  # name, workload, env...
    local name=$1 wl=$2; shift 2
    local line
    line=$(env "$@" timeout 100 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/ws_sweep_err.log | tail -1)
    echo "{\"name\": \"$name\", \"workload\": \"$wl\", \"line\": ${line:-null}}" >> $out
    python - "$name" "$wl" <<PY
import json,sys
try:
    l=json.loads('''$line''')
    print(sys.argv[1], sys.argv[2], "compress_ms=%.4f min=%.4f decompress_ms=%.4f frac=%.3f value=%.0f" % (l["compress_ms"], l["compress_ms_min"], l["decompress_ms"], l["roofline"]["frac"], l["value"]))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e)
PY
}
for wl in ${WORKLOADS:-cfg2}; do
    run v1 $wl NDZB_COMPRESS_KERNEL=v1
    for v in ${VARIANTS:-0 1 2 3 4}; do
        run ws$v $wl NDZB_WS_VARIANT=$v
    done
done
