"""Timeline of one pipelined offloader compress / decompress call (NDZB_PIPE_TRACE=1), 512^3 float, pinned buffers."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import ndzip_b200 as nz  # noqa: E402
from bench import make_device_input  # noqa: E402

shape, dtype = (512, 512, 512), "float32"
d = make_device_input(dtype, shape, device="cuda")
h = torch.empty(shape, dtype=d.dtype, pin_memory=True)
h.copy_(d)
h2 = torch.empty(shape, dtype=d.dtype, pin_memory=True)
h_stream = torch.empty(nz.compressed_length_bound(dtype, shape), dtype=torch.int32, pin_memory=True)
off = nz.make_cuda_offloader(dtype, 3)
for _ in range(3):
    n = off.compress(h, shape, h_stream)
    off.decompress(h_stream, n, h2, shape)
torch.cuda.synchronize()
os.environ["NDZB_PIPE_TRACE"] = "1"
for _ in range(2):
    t0 = time.perf_counter()
    n = off.compress(h, shape, h_stream)
    t1 = time.perf_counter()
    off.decompress(h_stream, n, h2, shape)
    t2 = time.perf_counter()
    print("wall: compress %.3f ms decompress %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
assert torch.equal(h.view(torch.int32), h2.view(torch.int32))
