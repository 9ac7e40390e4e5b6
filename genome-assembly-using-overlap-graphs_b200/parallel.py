"""The one exchange step of the sharded builder: gather the per-rank edge slices.

Candidate pairs are split into `world` contiguous slices of the global (a, b)-ordered pair
list (engine.candidate_pairs(shard=...)); each rank runs the DP and the edge expansion on its
slice with no communication.  Because the slices are contiguous, concatenating the rank
results in rank order IS the single-GPU edge list -- no merge.  This module moves the slices
to one rank with torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def gather_edges(edges: torch.Tensor, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """edges: int32[E_rank, 4] on every rank.  Returns the concatenated int32[E, 4] on `dst`
    (None elsewhere).  Sizes are exchanged first (one all_gather of an int64 per rank), then
    every rank sends its rows straight into its slot of the destination buffer."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return edges
    n_local = torch.tensor([edges.shape[0]], dtype=torch.int64, device=edges.device)
    sizes = torch.zeros(world, dtype=torch.int64, device=edges.device)
    dist.all_gather_into_tensor(sizes, n_local, group=group)
    sizes_h = sizes.cpu().tolist()
    flat = edges.contiguous().view(-1)
    if rank == dst:
        total = sum(sizes_h)
        out = torch.empty(total * 4, dtype=torch.int32, device=edges.device)
        ops = []
        off = 0
        for r, n in enumerate(sizes_h):
            seg = out[off * 4:(off + n) * 4]
            if r == dst:
                seg.copy_(flat)
            elif n:
                ops.append(dist.P2POp(dist.irecv, seg, r, group))
            off += n
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out.view(total, 4)
    if flat.numel():
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, flat, dst, group)]):
            w.wait()
    return None


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_cpus(device_index: int):
    """(numa node, CPUs of that node) the GPU hangs off, from sysfs; None when the platform does not say."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        return (node, cpus) if cpus else None
    except Exception:
        return None


class SharedEdgeSink:
    """Host-side landing zone for the edge rows of all ranks of ONE node: a POSIX shared-memory
    segment that every rank maps and page-locks (cudaHostRegister), so each GPU copies its slice
    over its own PCIe link straight to its final position and rank 0 ends up with the complete,
    ordered list in host memory -- no NVLink gather and no single-link funnel.

    Use as the `host_sink` of OverlapEngine.overlap_edges(): calling it with the local row count
    all-gathers the counts, (re)sizes the segment collectively if needed and returns the pinned
    CPU tensor slice this rank must fill.  After a barrier, `rows()` on rank 0 is the whole list.
    """

    def __init__(self, group=None, cuda: bool = True, initial_rows: int = 1 << 20, numa_local: bool = True):
        # numa_local: before the segment is page-locked every rank first-touches ITS share of it from a CPU of the
        # NUMA node its GPU hangs off, so the rank's device->host copies land in memory behind its own PCIe root
        # instead of crossing the socket interconnect to wherever rank 0's registration would have faulted the pages in
        self.numa_local = numa_local
        self.numa_node = None
        self.group = group
        self.cuda = cuda and torch.cuda.is_available()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.shm = None
        self.tensor = None
        self.capacity = 0
        self.total = 0
        self._registered = None
        self._ensure(initial_rows)

    def _release(self):
        if self._registered is not None:
            torch.cuda.cudart().cudaHostUnregister(self._registered)
            self._registered = None
        self.tensor = None
        if self.shm is not None:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
            self.shm = None

    def _ensure(self, rows: int):
        """Collective: every rank calls with the same `rows`."""
        from multiprocessing import shared_memory
        import numpy as np
        if rows <= self.capacity:
            return
        self._release()
        rows = int(rows + rows // 16 + 1024)
        rows = (rows + 255) // 256 * 256           # 4 KiB multiple: cudaHostRegister wants page-granular sizes
        nbytes = rows * 16
        name = [None]
        if self.rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self.shm.name
        dist.broadcast_object_list(name, src=0, group=self.group)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:                                   # the creator owns the segment: do not let this process unlink it
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        arr = np.ndarray((rows, 4), dtype=np.int32, buffer=self.shm.buf)
        self.tensor = torch.from_numpy(arr)
        if self.cuda and self.numa_local:
            self._first_touch(arr, rows)
            dist.barrier(group=self.group)
        if self.cuda:
            # page-lock the mapping in every process, one rank at a time: concurrent registration of
            # the same not-yet-populated segment fails with cudaErrorOperatingSystem (measured)
            for r in range(self.world):
                if r == self.rank:
                    err = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)
                    if int(err) != 0:
                        raise RuntimeError(f"cudaHostRegister of the shared edge buffer ({nbytes} bytes) failed: "
                                           f"cudaError {int(err)}")
                    self._registered = self.tensor.data_ptr()
                dist.barrier(group=self.group)
        self.capacity = rows
        dist.barrier(group=self.group)

    def _first_touch(self, arr, rows: int):
        """Fault in rows [rank, rank+1) * rows / world (the pair-range shards produce nearly equal edge counts)
        from a CPU next to this rank's GPU.  Pages of a fresh POSIX segment are placed on the node of the CPU that
        first writes them; the later cudaHostRegister pins them where they are."""
        import os
        where = gpu_numa_cpus(torch.cuda.current_device())
        lo, hi = rows * self.rank // self.world, rows * (self.rank + 1) // self.world
        old = None
        try:
            if where is not None:
                old = os.sched_getaffinity(0)
                cpus = where[1] & old
                if cpus:
                    os.sched_setaffinity(0, cpus)
                    self.numa_node = where[0]
            arr[lo:hi] = 0
        finally:
            if old is not None:
                os.sched_setaffinity(0, old)

    def __call__(self, n_local: int, e_begin: Optional[int] = None, total: Optional[int] = None) -> torch.Tensor:
        """The slice this rank must fill.  With (e_begin, total) -- the global row offset of the rank's slice
        and the global row count, both known to every rank from the join index -- no communication is
        needed (a re-size is collective, but every rank sees the same `total` and takes it together);
        otherwise the row counts are exchanged (one all_gather of an int64 per rank)."""
        if e_begin is None or total is None:
            sizes = torch.zeros(self.world, dtype=torch.int64)
            mine = torch.tensor([int(n_local)], dtype=torch.int64)
            if self.cuda and dist.get_backend(self.group) == "nccl":
                dev = torch.device("cuda", torch.cuda.current_device())
                sizes_d = sizes.to(dev)
                dist.all_gather_into_tensor(sizes_d, mine.to(dev), group=self.group)
                sizes = sizes_d.cpu()
            else:
                dist.all_gather_into_tensor(sizes, mine, group=self.group)
            sizes = sizes.tolist()
            total = int(sum(sizes))
            e_begin = int(sum(sizes[:self.rank]))
        self.total = int(total)
        self._ensure(self.total)
        return self.tensor[int(e_begin):int(e_begin) + int(n_local)]

    def rows(self):
        """NumPy view of the complete list (meaningful on every rank after a barrier)."""
        return self.tensor[:self.total].numpy()

    def close(self):
        dist.barrier(group=self.group)
        self._release()


class PeerEdgeBuffer:
    """Fused compute + gather over NVLink peer memory: the destination rank's edge buffer is a
    symmetric-memory allocation mapped into every rank's address space, and the DP kernel's fused
    epilogue (csrc/dp.cuh, DpEdgeOut) stores its 16-byte edge rows straight into it at the rank's
    global row offset.  No separate gather pass, no staging copy: the only collective left is the
    exchange of the per-rank row counts (one all-gather of an int64) and a barrier.

    `slot(n_local)` is the `sink` of OverlapEngine.overlap_edges_fused(): it returns the raw device
    address (in the destination's buffer) where this rank's rows start.
    """

    def __init__(self, rows: int, device, group=None, dst: int = 0):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.dst = dst
        self.capacity = int(rows + rows // 16 + 1024)
        self.local = symm.empty(self.capacity * 4, dtype=torch.int32, device=device)
        self.hdl = symm.rendezvous(self.local, self.group)
        self.dst_ptr = int(self.hdl.buffer_ptrs[dst])
        self.total = 0
        self._sizes = torch.zeros(self.world, dtype=torch.int64, device=device)

    def slot(self, n_local: int, e_begin: Optional[int] = None, total: Optional[int] = None) -> int:
        """Device address (inside the destination rank's buffer) of this rank's first row.  With
        (e_begin, total) from the join index no communication is needed; otherwise the row counts are
        exchanged first."""
        if e_begin is None or total is None:
            n = torch.tensor([int(n_local)], dtype=torch.int64, device=self.local.device)
            dist.all_gather_into_tensor(self._sizes, n, group=self.group)
            sizes = self._sizes.cpu().tolist()
            total = int(sum(sizes))
            e_begin = int(sum(sizes[:self.rank]))
        self.total = int(total)
        if self.total > self.capacity:
            raise RuntimeError(f"peer edge buffer too small: {self.total} rows > capacity {self.capacity}")
        return self.dst_ptr + 16 * int(e_begin)

    def barrier(self) -> None:
        """All ranks' kernels have finished storing (stream-ordered) and the stores are visible."""
        torch.cuda.current_stream(self.local.device).synchronize()
        self.hdl.barrier()

    def result(self):
        """The complete ordered edge list on the destination rank (None elsewhere)."""
        if self.rank != self.dst:
            return None
        return self.local[:self.total * 4].view(self.total, 4)
