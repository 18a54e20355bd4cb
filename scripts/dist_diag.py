#!/usr/bin/env python
"""Times the pieces of one multi-GPU step (ndzb_dist_*) separately, under torchrun: compress alone, decompress alone,
compress + decompress without joining the exchange, the full step. usage: torchrun ... scripts/dist_diag.py cfg5 [reps]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ndzip_b200 import dist as nzd  # noqa: E402
from bench import WORKLOADS, make_device_input  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    for name in sys.argv[1].split(","):
        dtype, shape, _ = WORKLOADS[name]
        tbits = torch.int32 if dtype == "float32" else torch.int64
        d_in = make_device_input(dtype, shape, device=dev, index_offset=rank * shape[0])
        codec = nzd.DistCodec(dtype, (shape[0] * world,) + tuple(shape[1:]))
        d_stream = torch.empty(int(codec.layout.local_bound_words), dtype=tbits, device=dev)
        d_len = torch.zeros(1, dtype=torch.int32, device=dev)
        d_back = torch.empty_like(d_in)

        def timed(fn, label):
            for _ in range(2):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
            tmax, tmin = t.clone(), t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"{name} world={world} {label:46s} max {tmax.item():.4f} ms  min {tmin.item():.4f} ms", flush=True)

        def full():
            codec.compress(d_in, d_stream, d_len)
            codec.decompress(d_stream, d_back)
            codec.wait_exchange()

        def no_join():
            codec.compress(d_in, d_stream, d_len)
            codec.decompress(d_stream, d_back)

        def compress_join():
            codec.compress(d_in, d_stream, d_len)
            codec.wait_exchange()

        timed(lambda: codec.compress(d_in, d_stream, d_len), "compress (exchange on the side stream)")
        timed(lambda: codec.decompress(d_stream, d_back), "decompress")
        timed(compress_join, "compress + join exchange")
        timed(no_join, "compress + decompress, exchange not joined")
        timed(full, "full step")
        codec.close()
        del d_in, d_stream, d_back
        torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
