"""Times the decompress launch alone (CUDA events) and checks the round trip. usage: dec_time.py workload [reps] [ENV=..,ENV=.. ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ndzip_b200 as nz
from bench import make_device_input, WORKLOADS

wl = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dtype, shape, _ = WORKLOADS[wl]
d_in = make_device_input(dtype, shape, device="cuda")
tbits = torch.int32 if dtype == "float32" else torch.int64
d_stream = torch.zeros(nz.compressed_length_bound(dtype, shape), dtype=tbits, device="cuda")
d_len = torch.zeros(1, dtype=torch.int32, device="cuda")
d_back = torch.empty_like(d_in)
nz.make_cuda_compressor(dtype, nz.compressor_requirements(shape)).compress(d_in, shape, d_stream, d_len)
torch.cuda.synchronize()
for env in sys.argv[3:] or [""]:
    for kv in env.split(","):
        if kv:
            k, v = kv.split("=")
            os.environ[k] = v
    dec = nz.make_cuda_decompressor(dtype, len(shape))
    ts = []
    for i in range(reps + 3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dec.decompress(d_stream, d_back, shape)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    ok = torch.equal(d_in.view(tbits), d_back.view(tbits))
    print("%s %-30s avg %.4f ms  min %.4f ms  round trip %s" % (wl, env, sum(ts) / len(ts), min(ts), "ok" if ok else "MISMATCH"), flush=True)
    for kv in env.split(","):
        if kv:
            os.environ.pop(kv.split("=")[0], None)
    del dec
