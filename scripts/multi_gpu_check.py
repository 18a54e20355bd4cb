#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
every rank compresses its slab, the ranks exchange compressed word counts over NCCL, headers are
fixed up, the stream is gathered to rank 0 over NVLink and compared bit for bit with the stream rank 0
gets by compressing the WHOLE grid on one GPU (which tests/ pin against the reference). Then every rank
decompresses its own local stream and checks the round trip.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ndzip_b200 as nz  # noqa: E402
from ndzip_b200 import dist as nzd  # noqa: E402
from bench import make_device_input  # noqa: E402

CASES = [("float32", (256, 128, 160)), ("float64", (4 * 64 + 17, 200)), ("float32", (9 * 4096 + 5,)), ("float64", (128, 64, 64))]


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ok = True
    for dtype, shape in CASES:
        tbits = torch.int32 if dtype == "float32" else torch.int64
        full = make_device_input(dtype, shape, seed=77, device=dev)  # identical on every rank (deterministic)
        b, e = nzd.slab_partition(shape, world)[rank]
        local_shape = nzd.slab_shape(shape, (b, e))
        slab = full[b:e].contiguous()
        comp = nz.make_cuda_compressor(dtype, local_shape)
        d_stream = torch.zeros(max(1, nz.compressed_length_bound(dtype, local_shape)), dtype=tbits, device=dev)
        d_len = torch.zeros(1, dtype=torch.int32, device=dev)
        comp.compress(slab, local_shape, d_stream, d_len)
        H = nz.num_hypercubes(local_shape)
        hdr = nzd.header_words(dtype, H)
        cube_words = d_len.to(torch.int64) - hdr - nzd.border_in(local_shape)
        layout = nzd.exchange_layout(dtype, shape, cube_words)
        # the exchange step as bench.py does it: all-gather the stream lengths, one fix-up kernel
        gathered = torch.zeros(world, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(gathered, d_len)
        spans = nzd.slab_partition(shape, world)
        overhead = torch.tensor([nzd.header_words(dtype, nzd.cubes_in(nzd.slab_shape(shape, sp))) + nzd.border_in(nzd.slab_shape(shape, sp))
                                 for sp in spans], dtype=torch.int32, device=dev)
        header32 = torch.zeros(max(H, 1), dtype=torch.int32, device=dev)
        if H:
            comp.fixup_header(d_stream[:hdr].view(torch.int32), header32, H, gathered, overhead, rank)
            check = d_stream[:hdr].view(torch.int32)[:H].clone()
            comp.add_offset(check, H, torch.tensor([layout.cube_word_base], dtype=torch.int32, device=dev))
            assert torch.equal(check, header32[:H]), "fixup_header and add_offset disagree"
        gathered = nzd.gather_global_stream(layout, d_stream, header32, root=0)
        # local round trip
        back = torch.empty_like(slab)
        nz.make_cuda_decompressor(dtype, len(shape)).decompress(d_stream, back, local_shape)
        torch.cuda.synchronize()
        rt = torch.equal(back.view(tbits), slab.view(tbits))
        if rank == 0:
            comp1 = nz.make_cuda_compressor(dtype, shape)
            ref_stream = torch.zeros(nz.compressed_length_bound(dtype, shape), dtype=tbits, device=dev)
            ref_len = torch.zeros(1, dtype=torch.int32, device=dev)
            comp1.compress(full, shape, ref_stream, ref_len)
            torch.cuda.synchronize()
            n = int(ref_len.item())
            same = n == gathered.numel() and torch.equal(gathered, ref_stream[:n])
            print(f"[multi-gpu x{world}] {dtype} {shape}: global stream {'IDENTICAL' if same else 'MISMATCH'} ({n} words), local round trip {'ok' if rt else 'FAILED'}")
            ok &= bool(same)
        ok &= bool(rt)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI-GPU PARITY", "PASS" if int(flag.item()) == 1 else "FAIL")
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
