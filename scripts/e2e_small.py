"""Offloader round trip on small and medium arrays (pinned buffers): GB/s per size, 1-D float."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import ndzip_b200 as nz  # noqa: E402
from bench import make_device_input  # noqa: E402

for mib in (16, 32, 64, 128, 256):
    shape = (mib << 18,)
    d = make_device_input("float32", shape, device="cuda")
    h = torch.empty(shape, dtype=d.dtype, pin_memory=True)
    h.copy_(d)
    h2 = torch.empty(shape, dtype=d.dtype, pin_memory=True)
    h_stream = torch.empty(nz.compressed_length_bound("float32", shape), dtype=torch.int32, pin_memory=True)
    off = nz.make_cuda_offloader("float32", 1)
    for _ in range(3):
        n = off.compress(h, shape, h_stream)
        off.decompress(h_stream, n, h2, shape)
    ts = []
    for _ in range(7):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = off.compress(h, shape, h_stream)
        t1 = time.perf_counter()
        off.decompress(h_stream, n, h2, shape)
        ts.append((time.perf_counter() - t0, t1 - t0))
    best = min(ts)
    assert torch.equal(h.view(torch.int32), h2.view(torch.int32))
    print("%4d MiB: compress %.3f ms decompress %.3f ms round trip %.1f GB/s" % (mib, best[1] * 1e3, (best[0] - best[1]) * 1e3, (mib << 20) / best[0] / 1e9), flush=True)
    del off, h, h2, h_stream, d
