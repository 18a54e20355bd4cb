#!/usr/bin/env python
"""Builds a variant of libndzip_b200.so under build/exp/ for A/B runs:  scripts/build_variant.py NAME [nvcc flags...]
Select it at run time with NDZB_LIB=build/exp/libndzb_NAME.so (ndzip_b200/_lib.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ndzip_b200 import build  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "build", "exp", f"libndzb_{name}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
print(build.build(force=True, out=out, extra_flags=flags))
