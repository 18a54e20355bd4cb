"""PCIe ceiling vs the offloader: pinned H2D / D2H bandwidth alone and together, then offloader compress / decompress."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ndzip_b200 as nz
from bench import make_device_input

shape, dtype = (512, 512, 512), "float32"
d = make_device_input(dtype, shape, device="cuda")
h = torch.empty(shape, dtype=d.dtype, pin_memory=True); h.copy_(d)
h2 = torch.empty(shape, dtype=d.dtype, pin_memory=True)
d2 = torch.empty_like(d)
nbytes = d.numel() * 4
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts)
t = timed(lambda: d.copy_(h, non_blocking=True)); print("H2D 512 MiB pinned: %.2f ms  %.1f GB/s" % (t * 1e3, nbytes / t / 1e9))
t = timed(lambda: h2.copy_(d, non_blocking=True)); print("D2H 512 MiB pinned: %.2f ms  %.1f GB/s" % (t * 1e3, nbytes / t / 1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
t = timed(both); print("H2D + D2H concurrently: %.2f ms  %.1f GB/s per direction" % (t * 1e3, nbytes / t / 1e9))
bound = nz.compressed_length_bound(dtype, shape)
h_stream = torch.empty(bound, dtype=torch.int32, pin_memory=True)
off = nz.make_cuda_offloader(dtype, 3)
n = off.compress(h, shape, h_stream); off.decompress(h_stream, n, h2, shape)
tc = timed(lambda: off.compress(h, shape, h_stream)); td = timed(lambda: off.decompress(h_stream, n, h2, shape))
sb = n * 4
print("offloader compress: %.2f ms (in %.1f GB/s, out %.0f MB)   decompress: %.2f ms (out %.1f GB/s)   round trip %.1f GB/s" % (
    tc * 1e3, nbytes / tc / 1e9, sb / 1e6, td * 1e3, nbytes / td / 1e9, nbytes / (tc + td) / 1e9))
for cb in (16, 32, 64, 128):
    os.environ["NDZB_CHUNK_BYTES"] = str(cb << 20)
    off2 = nz.make_cuda_offloader(dtype, 3)
    off2.compress(h, shape, h_stream); off2.decompress(h_stream, n, h2, shape)
    tc = timed(lambda: off2.compress(h, shape, h_stream)); td = timed(lambda: off2.decompress(h_stream, n, h2, shape))
    print("chunk %3d MiB: compress %.2f ms decompress %.2f ms round trip %.1f GB/s" % (cb, tc * 1e3, td * 1e3, nbytes / (tc + td) / 1e9))
