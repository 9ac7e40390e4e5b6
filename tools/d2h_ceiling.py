"""Aggregate device->host bandwidth of N ranks copying at once (what bounds the N-GPU e2e leg: all edge rows must
land in host memory).  torchrun --nproc-per-node N tools/d2h_ceiling.py
  (a) every rank into its own cudaHostAlloc buffer; (b) every rank into its slice of ONE shared, page-locked segment
  (parallel.SharedEdgeSink -- the e2e path), faulted in by rank 0's registration; (c) the same with every rank's share
  first-touched from a CPU of its GPU's NUMA node."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

PKG = "genome-assembly-using-overlap-graphs_b200"
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
par = importlib.import_module(PKG + ".parallel")
rows = (2 << 30) // 16                         # 2 GiB of 16-byte rows per rank
dev = torch.empty((rows, 4), dtype=torch.int32, device="cuda").fill_(rank)
own = torch.empty((rows, 4), dtype=torch.int32).pin_memory()


def timed(dst, reps=3):
    best = 0.0
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        dst.copy_(dev, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        best = max(best, world * rows * 16 / dt / 1e9)
    return best


out = {"ranks": world, "bytes_per_rank": rows * 16}
out["own_pinned_buffers_gbs"] = timed(own)
out["gpu_numa"] = par.gpu_numa_cpus(lr)[0] if par.gpu_numa_cpus(lr) else None
if world > 1:
    nodes = [None] * world
    dist.all_gather_object(nodes, out["gpu_numa"])
    out["gpu_numa"] = nodes
    for numa in (False, True):
        sink = par.SharedEdgeSink(initial_rows=rows * world, numa_local=numa)
        mine = sink(rows, rank * rows, rows * world)
        out["shared_segment_numa_local_gbs" if numa else "shared_segment_gbs"] = timed(mine)
        sink.close()
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
