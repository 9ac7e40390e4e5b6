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
