"""Multi-GPU parity on hardware (needs >= 2 GPUs; skipped otherwise): every exchange path of the sharded
builder -- the NCCL gather, the DP epilogue's peer-memory stores (PeerEdgeBuffer) and the shared
page-locked host sink -- must reproduce the single-GPU edge list byte for byte, on a workload with
duplicate reads (BASELINE.json configs[1]: PhiX-like, N = 50,000)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, load_pkg, has_cuda


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not has_cuda() or _n_gpus() < 2, reason="needs at least two CUDA devices")]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    report = {}
    try:
        synth = load_pkg("synth")
        par = load_pkg("parallel")
        eng = load_pkg("engine").get_engine(rank)
        bases, offsets = synth.make_workload("phix_n50000_l150")
        ub, uo, counts, _ = synth.dedup(bases, offsets)
        assert counts.max() > 1                                   # the copy x copy expansion is exercised
        U = len(counts)
        node_off = np.zeros(U + 1, np.int64)
        np.cumsum(counts, out=node_off[1:])
        k = 5
        rs = eng.upload_reads(ub, uo)
        d_copies = torch.from_numpy(counts).to(eng.device)
        d_node_off = torch.from_numpy(node_off).to(eng.device)
        idx = eng.kmer_index(rs, k)

        # the single-GPU list, computed by every rank on its own GPU
        pa, pb, _ = eng.candidate_pairs(rs, idx, k)
        whole = eng.overlap_edges_fused(rs, pa, pb, d_copies, d_node_off)
        E = int(whole.shape[0])
        whole_hash = int(eng.edge_hash(whole).item())

        # this rank's shard
        sa, sb, p_begin = eng.candidate_pairs(rs, idx, k, (rank, world))
        assert p_begin == int(pa.shape[0]) * rank // world
        mine = eng.overlap_edges_fused(rs, sa, sb, d_copies, d_node_off)

        # (1) NCCL send/recv gather
        got = par.gather_edges(mine, 0)
        if rank == 0:
            report["gather_equal"] = bool(torch.equal(got, whole))
            report["gather_hash_equal"] = int(eng.edge_hash(got).item()) == whole_hash
        else:
            assert got is None
        # shard hashes with their global row offsets add up to the whole list's hash
        sizes = torch.zeros(world, dtype=torch.int64, device=eng.device)
        dist.all_gather_into_tensor(sizes, torch.tensor([mine.shape[0]], dtype=torch.int64, device=eng.device))
        first = int(sizes[:rank].sum().item())
        part = eng.edge_hash(mine, first_row=first)
        dist.all_reduce(part)                                     # wrapping int64 sum == u64 sum mod 2^64
        report["shard_hashes_add_up"] = int(part.item()) == whole_hash

        # (2) DP epilogue stores rows into rank 0's buffer over NVLink peer memory
        peer = par.PeerEdgeBuffer(E, eng.device)
        out = eng.overlap_edges_fused(rs, sa, sb, d_copies, d_node_off, sink=peer.slot)
        assert out is None
        peer.barrier()
        res = peer.result()
        if rank == 0:
            report["peer_equal"] = bool(torch.equal(res, whole))
        else:
            assert res is None
        peer.barrier()

        # (3) host result: every rank copies its slice into one shared page-locked host buffer
        sink = par.SharedEdgeSink(initial_rows=1024)             # small: forces the collective re-size
        eng.overlap_edges(ub, uo, counts, k, (rank, world), host_sink=sink)
        dist.barrier()
        if rank == 0:
            report["sink_equal"] = bool(np.array_equal(sink.rows(), whole.cpu().numpy()))
            report["sink_hash_equal"] = (eng.edge_hash_host(sink.rows()) & 0xFFFFFFFFFFFFFFFF) == (whole_hash & 0xFFFFFFFFFFFFFFFF)
        dist.barrier()
        # (4) the same with the inputs uploaded once in total (each rank 1/world + an all-gather over NVLink)
        sink.rows()[:] = 0
        dist.barrier()
        eng.overlap_edges(ub, uo, counts.astype(np.int32), k, (rank, world), host_sink=sink, upload_group=dist.group.WORLD)
        dist.barrier()
        if rank == 0:
            report["sharded_upload_equal"] = bool(np.array_equal(sink.rows(), whole.cpu().numpy()))
        d_whole = eng._to_device_sharded(ub, torch.uint8, dist.group.WORLD, slack=64)
        report["sharded_upload_bytes_equal"] = bool(np.array_equal(d_whole[:len(ub)].cpu().numpy(), ub))
        sink.close()
        report["edges"] = E
        report["ok"] = True
    except Exception as exc:                                      # noqa: BLE001
        import traceback
        report["ok"] = False
        report["error"] = f"rank {rank}: {type(exc).__name__}: {exc}\n{traceback.format_exc()}"
    q.put((rank, report))
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:                                             # noqa: BLE001
        pass


def test_two_gpu_exchange_paths_equal_single_gpu():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    reports = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    for r in range(world):
        assert reports[r]["ok"], reports[r].get("error")
        assert reports[r]["shard_hashes_add_up"]
    r0 = reports[0]
    assert r0["edges"] > 1_000_000
    assert r0["gather_equal"] and r0["gather_hash_equal"]
    assert r0["peer_equal"]
    assert r0["sink_equal"] and r0["sink_hash_equal"]
    assert r0["sharded_upload_equal"]
    assert all(reports[r]["sharded_upload_bytes_equal"] for r in range(world))
    for p in procs:
        assert p.exitcode == 0
