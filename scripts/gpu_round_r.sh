#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cli.py -m gpu -q --timeout 300 -p no:cacheprovider --tb=short > gpurun_out/r_pytest_cli.log 2>&1; echo "[cli tests] rc=$? $(tail -1 gpurun_out/r_pytest_cli.log)"
# file-level throughput of the tool: 512^3 float (512 MiB) from / to tmpfs
python - <<'PY'
import sys, os, time, subprocess
sys.path.insert(0, os.getcwd())
import torch
from bench import make_device_input
d = make_device_input("float32", (512, 512, 512), device="cuda").cpu().numpy()
d.tofile("/dev/shm/nz_in.bin")
tool = "ndzip_b200/bin/ndzip-compress"
for mode in ([], ["--no-mmap"]):
    t0 = time.perf_counter(); r = subprocess.run([tool, "-n", "512", "512", "512", "-i", "/dev/shm/nz_in.bin", "-o", "/dev/shm/nz_out.ndz", *mode], stderr=subprocess.PIPE); t1 = time.perf_counter()
    r2 = subprocess.run([tool, "-d", "-n", "512", "512", "512", "-i", "/dev/shm/nz_out.ndz", "-o", "/dev/shm/nz_back.bin", *mode], stderr=subprocess.PIPE); t2 = time.perf_counter()
    same = open("/dev/shm/nz_back.bin", "rb").read() == open("/dev/shm/nz_in.bin", "rb").read()
    print("cli %-9s compress %.2f s (%.2f GB/s incl. process start + CUDA init) decompress %.2f s  round trip %s | %s" % (
        " ".join(mode) or "mmap", t1 - t0, d.nbytes / (t1 - t0) / 1e9, t2 - t1, "identical" if same else "MISMATCH", r.stderr.decode().strip()))
for f in ("nz_in.bin", "nz_out.ndz", "nz_back.bin"):
    os.remove("/dev/shm/" + f)
PY
