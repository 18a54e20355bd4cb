#!/usr/bin/env python
"""Prints the headline numbers of bench.py JSON lines given as files."""
import json
import sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    r = d.get("roofline", {})
    print(f"{f}: {d['config']['workload'][:40]:40s} n={d['n_gpus']} value={d['value']:.0f} c={d.get('compress_gbs', 0):.0f} GB/s ({d.get('compress_ms', 0):.3f} ms) "
          f"d={d.get('decompress_gbs', 0):.0f} GB/s ({d.get('decompress_ms', 0):.3f} ms) ratio={d['config'].get('ratio', 0):.3f} "
          f"frac={r.get('frac', 0):.3f} dfrac={r.get('decompress_kernel', {}).get('frac', 0):.3f} e2e={d.get('e2e', {}).get('value', 0):.1f} "
          f"cpu={d.get('cpu_baseline', {}).get('value', 0):.1f} clk={d.get('clocks', {}).get('sm_mhz') if d.get('clocks') else None}")
