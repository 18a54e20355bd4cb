"""Smallest run of compress_ws_kernel: a few cubes, checked against the oracle (used under compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ndzip_b200 as nz
from ndzip_b200 import synth
from oracle import get_oracle

dtype = sys.argv[1] if len(sys.argv) > 1 else "float32"
shape = tuple(int(x) for x in sys.argv[2].split("x")) if len(sys.argv) > 2 else (20 * 4096,)
gen = sys.argv[3] if len(sys.argv) > 3 else "smooth"
data = synth.make(gen, shape, dtype, seed=3)
tbits = torch.int32 if dtype == "float32" else torch.int64
d_in = torch.from_numpy(data).cuda()
d_stream = torch.zeros(nz.compressed_length_bound(dtype, shape), dtype=tbits, device="cuda")
d_len = torch.zeros(1, dtype=torch.int32, device="cuda")
comp = nz.make_cuda_compressor(dtype, shape)
comp.compress(d_in, shape, d_stream, d_len)
torch.cuda.synchronize()
n = int(d_len.item())
expect = get_oracle().compress(data)
got = d_stream[:n].cpu().numpy().view(expect.dtype)
print("words", n, "expect", expect.size, "equal", n == expect.size and np.array_equal(got, expect))
