"""8-GPU host placement probe (torchrun, one rank per GPU): topology, then pinned H2D || D2H bandwidth with all ranks
copying at once — first with the buffers wherever the process happened to run, then after ndzb_bind_host_to_device."""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import ndzip_b200 as nz  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")


def barrier():
    if world > 1:
        dist.barrier()


if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["sh", "-c", "cat /sys/devices/system/node/node*/cpulist; nproc; free -g | head -2"]):
        try:
            print(subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout[:3000], flush=True)
        except Exception as e:  # noqa: BLE001
            print(cmd, e)
barrier()
print("rank %d gpu %d bus %s numa node %d affinity %d cpus" % (
    rank, local, torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else "?",
    nz.device_numa_node(local), len(os.sched_getaffinity(0))), flush=True)

nbytes = 512 << 20
d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def measure(tag):
    h_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_a.fill_(1)
    h_b.fill_(2)
    res = {}
    for what in ("h2d", "d2h", "both"):
        ts = []
        for _ in range(4):
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            if what in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if what in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        res[what] = nbytes / min(ts[1:]) / 1e9
    print("%s rank %d: H2D %.1f GB/s  D2H %.1f GB/s  both %.1f GB/s per direction" % (tag, rank, res["h2d"], res["d2h"], res["both"]), flush=True)
    return h_a, h_b


measure("unbound")
barrier()
node = nz.bind_host_to_device(local)
print("rank %d bound to node %d, affinity now %d cpus" % (rank, node, len(os.sched_getaffinity(0))), flush=True)
barrier()
measure("bound  ")
barrier()
# the offloader round trip from bound buffers, all ranks at once
shape = (512, 512, 512)
from bench import make_device_input  # noqa: E402
d = make_device_input("float32", shape, device="cuda")
h = torch.empty(shape, dtype=d.dtype, pin_memory=True)
h.copy_(d)
h2 = torch.empty(shape, dtype=d.dtype, pin_memory=True)
h_stream = torch.empty(nz.compressed_length_bound("float32", shape), dtype=torch.int32, pin_memory=True)
off = nz.make_cuda_offloader("float32", 3)
n = off.compress(h, shape, h_stream)
off.decompress(h_stream, n, h2, shape)
ts = []
for _ in range(4):
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    n = off.compress(h, shape, h_stream)
    off.decompress(h_stream, n, h2, shape)
    ts.append(time.perf_counter() - t0)
print("bound   rank %d: offloader round trip %.2f ms = %.1f GB/s" % (rank, min(ts) * 1e3, d.numel() * 4 / min(ts) / 1e9), flush=True)
if world > 1:
    dist.destroy_process_group()
