#!/bin/bash
# Round evidence in one GPU call: ncu --set full captures of the compress and decompress kernels for BASELINE configs
# 2-5 (one launch each), the launch list of a bench run, and the bench lines. Output under gpurun_out/<tag>_*.
# usage: scripts/profile_round.sh r2 "cfg2 cfg3" [bench]; summaries are made afterwards with scripts/ncu_summary.py (no
# GPU needed). gpurun copies back at most 64 MiB: two workloads (four ~9 MB reports) per call.
tag=${1:-r2}
wls=${2:-"cfg2 cfg3"}
bench=${3:-}
mkdir -p gpurun_out
common="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-sustained"
for wl in $wls; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:compress_ws -s 4 -c 1 -f -o gpurun_out/${tag}_${wl}_compress_ws \
      python bench.py --workload $wl $common > gpurun_out/${tag}_${wl}_ncu_c.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:decompress -s 4 -c 1 -f -o gpurun_out/${tag}_${wl}_decompress \
      python bench.py --workload $wl $common > gpurun_out/${tag}_${wl}_ncu_d.log 2>&1
done
[ -z "$bench" ] && exit 0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:compress|decompress|border|fixup" -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-sustained > gpurun_out/${tag}_launches.log 2>&1
for wl in cfg2 cfg1 cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/${tag}_bench_${wl}.json 2> gpurun_out/${tag}_bench_${wl}.err
  python scripts/bench_summary.py gpurun_out/${tag}_bench_${wl}.json
done
ls -la gpurun_out | tail -30
