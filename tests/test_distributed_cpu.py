"""The N>1 exchange step on CPU: world_size-2 gloo (host logic only, no GPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_pkg, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sizes, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = load_pkg("parallel")
    start = sum(sizes[:rank])
    rows = torch.arange(start * 4, (start + sizes[rank]) * 4, dtype=torch.int32).view(-1, 4)
    out = par.gather_edges(rows, dst=0)
    if rank == 0:
        q.put(out.numpy().copy())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def _run(sizes):
    world = len(sizes)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sizes, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    total = sum(sizes)
    assert np.array_equal(got, np.arange(total * 4, dtype=np.int32).reshape(total, 4))


def _sink_worker(rank, world, port, sizes, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = load_pkg("parallel")
    sink = par.SharedEdgeSink(cuda=False, initial_rows=2)
    for rnd in range(2):                                     # second round forces a collective re-size
        n = sizes[rank] * (1 + 300 * rnd)
        all_n = [s * (1 + 300 * rnd) for s in sizes]
        start = sum(all_n[:rank])
        dst = sink(n)
        dst.copy_(torch.arange(start * 4, (start + n) * 4, dtype=torch.int32).view(-1, 4))
        dist.barrier()
        if rank == 0:
            q.put(sink.rows().copy())
        dist.barrier()
    sink.close()
    dist.destroy_process_group()


def test_shared_edge_sink_world2():
    sizes = [5, 3]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sink_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    for rnd in range(2):
        got = q.get(timeout=120)
        total = sum(sizes) * (1 + 300 * rnd)
        assert np.array_equal(got, np.arange(total * 4, dtype=np.int32).reshape(total, 4))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0


def test_gather_edges_world2():
    _run([5, 3])


def test_gather_edges_world2_with_empty_rank():
    _run([0, 7])
    _run([4, 0])
