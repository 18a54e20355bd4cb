"""Times the compress and the decompress launch alone (CUDA events) for one or more BASELINE workloads and checks
the round trip. A/B runs: NDZB_LIB=build/exp/libndzb_X.so python scripts/kernel_time.py cfg2,cfg3 [reps] [ENV=V,...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import ndzip_b200 as nz  # noqa: E402
from bench import WORKLOADS, make_device_input, measured_hbm_peak  # noqa: E402

reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for kv in (sys.argv[3].split(",") if len(sys.argv) > 3 else []):
    k, v = kv.split("=")
    os.environ[k] = v
peak, _ = measured_hbm_peak()
tag = os.path.basename(os.environ.get("NDZB_LIB", "default"))
for wl in sys.argv[1].split(","):
    dtype, shape, _ = WORKLOADS[wl]
    d_in = make_device_input(dtype, shape, device="cuda")
    tbits = torch.int32 if dtype == "float32" else torch.int64
    d_stream = torch.zeros(nz.compressed_length_bound(dtype, shape), dtype=tbits, device="cuda")
    d_len = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_back = torch.zeros_like(d_in)
    comp = nz.make_cuda_compressor(dtype, nz.compressor_requirements(shape))
    dec = nz.make_cuda_decompressor(dtype, len(shape))

    def timed(fn):
        ts = []
        for i in range(reps + 3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        return sum(ts) / len(ts), min(ts)

    tc = timed(lambda: comp.compress(d_in, shape, d_stream, d_len))
    td = timed(lambda: dec.decompress(d_stream, d_back, shape))
    ok = torch.equal(d_in.view(tbits), d_back.view(tbits))
    nbytes = d_in.numel() * d_in.element_size()
    algo = nbytes + int(d_len.item()) * d_in.element_size()
    print("%-22s %s compress avg %.4f min %.4f ms frac %.3f | decompress avg %.4f min %.4f ms frac %.3f | roundtrip %s" % (
        tag, wl, tc[0], tc[1], algo / (tc[0] * 1e-3) / 1e9 / peak, td[0], td[1], algo / (td[0] * 1e-3) / 1e9 / peak,
        "ok" if ok else "MISMATCH"), flush=True)
    del comp, dec, d_in, d_stream, d_back
