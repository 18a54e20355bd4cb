#!/usr/bin/env python
"""Multi-GPU parity check of the library's data plane (ndzb_dist_*; run under torchrun, one rank per GPU):
every rank compresses its slab, the library exchanges the stream lengths over NCCL and fixes the headers up, gathers
the stream on rank 0 and rank 0 compares it bit for bit with the stream it gets by compressing the WHOLE grid on one
GPU (which tests/ pin against the reference). Then every rank decompresses its own local stream and checks the round
trip. Shapes with borders in every dimension, odd cube counts (f64 header padding word) and empty slabs.

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

CASES = [("float32", (256, 128, 160)), ("float64", (4 * 64 + 17, 200)), ("float32", (9 * 4096 + 5,)), ("float64", (128, 64, 64)),
         ("float64", (3 * 16 + 5, 16, 50)), ("float32", (20, 40, 40)), ("float64", (7 * 4096,))]


def check_case(dtype, shape, rank, world, dev):
    tbits = torch.int32 if dtype == "float32" else torch.int64
    full = make_device_input(dtype, shape, seed=77, device=dev)  # identical on every rank (deterministic)
    codec = nzd.DistCodec(dtype, shape)
    L = codec.layout
    slab = full[L.slab_begin:L.slab_end].contiguous()
    d_stream = torch.zeros(max(1, int(L.local_bound_words)), dtype=tbits, device=dev)
    d_len = torch.zeros(1, dtype=torch.int32, device=dev)
    codec.compress(slab, d_stream, d_len)
    d_global = torch.zeros(max(1, int(L.global_bound_words)), dtype=tbits, device=dev) if rank == 0 else None
    total = codec.gather(d_stream, d_global, root=0)
    path = codec.last_gather_path
    back = torch.empty_like(slab)
    codec.decompress(d_stream, back)
    torch.cuda.synchronize()
    ok = bool(torch.equal(back.view(tbits), slab.view(tbits)))
    if rank == 0:
        comp = nz.make_cuda_compressor(dtype, shape)
        ref = torch.zeros(max(1, nz.compressed_length_bound(dtype, shape)), dtype=tbits, device=dev)
        ref_len = torch.zeros(1, dtype=torch.int32, device=dev)
        comp.compress(full, shape, ref, ref_len)
        torch.cuda.synchronize()
        n = int(ref_len.cpu().numpy().view(np.uint32)[0])
        same = n == total and bool(torch.equal(ref[:n], d_global[:n]))
        ok = ok and same
        print(f"{dtype} {shape}: global stream {total} words, single-GPU {n} words, identical={same}, gather via {path}", flush=True)
    # ---- sharded container: no offset exchange, no gather. Every rank writes its own slab stream into
    # one file (ndzb_container_create_file / _write_segment), reads the table back, decodes ITS segment from the
    # container with ndzb_container_decompress_segment, and rank 0 converts the container to the single stream
    # (ndzb_container_to_global_stream), which must be the gathered one.
    n_local = int(d_len.cpu().numpy().view(np.uint32)[0])
    box = [None]
    if rank == 0:
        import tempfile
        box[0] = os.path.join(tempfile.gettempdir(), f"ndzb_check_{os.getpid()}.ndzs")
    dist.broadcast_object_list(box, src=0)
    path = box[0]
    hdr = nzd.write_sharded(path, dtype, shape, d_stream[:n_local])
    blob = np.fromfile(path, dtype=np.uint8)
    hdr2, mine = nzd.read_sharded(path, rank, world)
    seg_ok = len(mine) == 1 and mine[0][0] == (L.slab_begin, L.slab_end) and np.array_equal(
        mine[0][2].view(np.uint8), d_stream[:n_local].cpu().numpy().view(np.uint8))
    h_slab = np.zeros(tuple(slab.shape), dtype=dtype)
    if slab.numel():
        off = nz.make_cuda_offloader(dtype, len(shape))
        nzd.decompress_segment(off, blob, rank, h_slab)
        off.close()
    seg_ok = seg_ok and h_slab.tobytes() == slab.cpu().numpy().tobytes() and hdr2.total_bytes == blob.size == hdr.total_bytes
    if rank == 0:
        stitched = nzd.to_global_stream(blob)
        same = stitched.size == total and np.array_equal(stitched.view(np.uint8), d_global[:total].cpu().numpy().view(np.uint8))
        seg_ok = seg_ok and same
        print(f"    container: {hdr.total_bytes} bytes in {len(hdr.segments)} segments, every rank decoded its own; "
              f"single stream from the container identical={same}", flush=True)
    dist.barrier()
    if rank == 0:
        os.remove(path)
    ok = ok and bool(seg_ok)
    codec.close()
    return ok


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ok = True
    for gather in ("", "nccl"):  # the gather over NVLink peer memory (default) and over ncclSend / ncclRecv
        if gather:
            os.environ["NDZB_GATHER"] = gather
        else:
            os.environ.pop("NDZB_GATHER", None)
        for dtype, shape in CASES:
            ok = check_case(dtype, shape, rank, world, dev) and ok
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL", f"world={world}", flush=True)
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
