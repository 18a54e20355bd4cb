#!/bin/bash
# Builds libndzip_b200.so from the csrc/ of another git revision into build/exp/libndzb_<name>.so (A/B runs on one GPU box:
# timings differ by a few percent between boxes, so both builds have to run in the same gpurun call).
# usage: scripts/build_rev.sh <git-rev> <name> [nvcc flags...]
set -e
rev=$1; name=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p "$tmp/ndzip_b200/csrc" "$tmp/include"
git -C "$root" archive "$rev" ndzip_b200/csrc include | tar -x -C "$tmp"
# the sources include ../../include/...: keep the relative layout
NDZB_CSRC="$tmp/ndzip_b200/csrc" python "$root/scripts/build_variant.py" "$name" "$@"
rm -rf "$tmp"
