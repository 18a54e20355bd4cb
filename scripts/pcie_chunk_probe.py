"""How much does cutting the two concurrent PCIe transfers of an offloader call into chunks cost? Pure pinned copies, no
kernels: H2D of `a` bytes in n pieces on one stream || D2H of `b` bytes in m pieces on another (512^3 float call sizes)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

big, small = 512 << 20, 306 << 20
h_in = torch.empty(big, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(big, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(big, dtype=torch.uint8, device="cuda")
d_b = torch.empty(big, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d_bytes, n, d2h_bytes, m, d2h_delay_chunks=0):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev = []
        with torch.cuda.stream(s1):
            step = h2d_bytes // n
            for i in range(n):
                d_a[i * step:(i + 1) * step].copy_(h_in[i * step:(i + 1) * step], non_blocking=True)
                e = torch.cuda.Event()
                e.record(s1)
                ev.append(e)
        with torch.cuda.stream(s2):
            step = d2h_bytes // m
            if d2h_delay_chunks:
                s2.wait_event(ev[d2h_delay_chunks - 1])
            for i in range(m):
                h_out[i * step:(i + 1) * step].copy_(d_b[i * step:(i + 1) * step], non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


for what, a, b in (("compress call: H2D 512 MiB || D2H 306 MiB", big, small), ("decompress call: H2D 306 MiB || D2H 512 MiB", small, big)):
    print(what)
    for n, m in ((1, 1), (4, 4), (8, 8), (16, 16), (32, 32), (1, 16), (16, 1), (4, 16), (16, 4)):
        print("  H2D in %2d pieces, D2H in %2d pieces: %.2f ms" % (n, m, run(a, n, b, m)), flush=True)
    print("  16 / 16, D2H starts after the first H2D piece: %.2f ms" % run(a, 16, b, 16, 1), flush=True)
