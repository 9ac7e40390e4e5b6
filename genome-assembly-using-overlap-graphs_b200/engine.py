"""Host orchestration of the overlap-detection pipeline on one B200.

Stages (each a call into libovl_b200.so, kernels in csrc/):
  K0 pack_reads -> K1 kmer_keys -> K2 index_build -> K3 join_count/fill -> K4/K5 overlap_dp
  -> K6 expand_edges
which together replace overlapGraphs.py:30-60 (index, candidate lookup, DP call, edge
expansion).  PyTorch is used for device buffers, streams and host<->device copies only.
"""
from __future__ import annotations

import ctypes
import threading
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as nat

INDEL_DEFAULT = -2 ** 31          # aligners.py:7


@dataclass
class ReadSet:
    """Unique reads resident in HBM: 2-bit packed rows + lengths."""
    packed: torch.Tensor          # uint8 view of uint32[U * row_words]
    length: torch.Tensor          # int32[U]
    bad: torch.Tensor             # int32[1]: words with a non-ACGT byte
    row_words: int
    n_reads: int
    max_len: int
    code_bits: int = 2            # 2: 2-bit packed A/C/G/T rows; 8: byte rows (any alphabet, general kernels)


@dataclass
class KmerIndex:
    k: int
    prefix_key: torch.Tensor      # int64 view of uint64[U]
    suffix_key: torch.Tensor
    sorted_key: torch.Tensor
    sorted_uid: torch.Tensor      # int32 view of uint32[U]
    n_indexed: torch.Tensor       # int64[1]
    key_bits: int = 0             # significant key bits (2k + read-set tag bits); 64 for hashed keys
    table: Optional[torch.Tensor] = None      # int32[2^table_bits + 1] direct-address bucket table, or None
    table_bits: int = 0
    pos_of: Optional[torch.Tensor] = None     # int32[U] sorted position of every indexed read, or None


@dataclass
class Candidates:
    """Result of the K0-K3 job (build_candidates): packed reads, index, per-source join arrays and the
    scalars the host read back in one sync."""
    rs: ReadSet
    index: KmerIndex
    bucket_lo: torch.Tensor       # int32[U]
    self_rank: torch.Tensor       # int32[U]
    pair_off: torch.Tensor        # int64[U+1]
    edge_base: Optional[torch.Tensor]     # int64[U+1] when reads have copies
    cum: Optional[torch.Tensor]           # int64[U+1] copies scanned along the sorted index
    copies: Optional[torch.Tensor]
    node_off: Optional[torch.Tensor]
    total_pairs: int
    total_edges: int
    p_begin: int                  # this rank's slice of the pair list
    p_end: int
    cut_pairs: list               # 65 pair indices cutting the slice into 64 equal parts ...
    cut_edges: list               # ... and the first edge row of each
    arena: torch.Tensor

    @property
    def e_begin(self) -> int:
        return self.cut_edges[0]

    @property
    def e_end(self) -> int:
        return self.cut_edges[-1]


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def d2h_chunk_cuts(P: int, chunk_pairs: int = 1 << 62, min_chunked_pairs: int = 1_000_000):
    """Chunk schedule of the device->host copy that runs behind the DP, as cut points in 64ths of the slice (the
    cut points the K0-K3 job computed): 4, 8, 13, 13, 9, 6, 4, 3, 2, 1, 1.

    Small chunks first so the copy engine starts early, then chunks that shrink by ~2/3 per step: every copy hides
    behind the next DP chunk as long as copying a slice takes less than ~2/3 of computing it, and only the last 1/64
    is exposed.  (Round 2 first used 1/2, 1/4, ... 1/64: fine while the copy costs < 1/2 of the DP -- one to four
    GPUs -- but with eight GPUs sharing the host's 92 GB/s the copy is 0.66 of the DP and a first chunk of 1/2 left
    50 ms of it exposed.)  chunk_pairs caps the chunk size for very long lists; short lists are not chunked."""
    if P < min_chunked_pairs:
        return [0, 64]
    cuts = [0, 4, 12, 25, 38, 47, 53, 57, 60, 62, 63, 64]
    per64 = max(1, P // 64)
    step = max(1, chunk_pairs // per64)               # largest chunk, in 64ths
    return [c for lo, hi in zip(cuts[:-1], cuts[1:]) for c in range(lo, hi, step)] + [64]


class OverlapEngine:
    def __init__(self, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device visible: the overlap path runs on sm_100a only (no CPU fallback)")
        if device is None:
            device = torch.cuda.current_device()
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        ctx = ctypes.c_void_p()
        nat.check(nat.lib.ovl_ctx_create(self.device_index, ctypes.byref(ctx)))
        self._ctx = ctx
        self.sm_count = int(nat.lib.ovl_ctx_sm_count(ctx))
        self._total_mem = int(torch.cuda.get_device_properties(self.device).total_memory)
        self._pinned_out = None    # reusable pinned host buffer for edge rows (D2H at full PCIe rate)
        # one engine per device is shared by the whole process and hands out views of ONE pinned result
        # buffer: callers that consume such a view (the graph builders) hold this lock while they do
        self.lock = threading.RLock()
        # the scalars of the K0-K3 job land here (page-locked, written by the device, read after one event wait)
        self._totals = torch.zeros(int(nat.lib.ovl_totals_len()), dtype=torch.int64).pin_memory()

    @property
    def launches(self) -> int:
        """Kernels launched through this engine's context so far (counted inside libovl_b200.so)."""
        return int(nat.lib.ovl_ctx_launch_count(self._ctx)) if getattr(self, "_ctx", None) else 0

    @staticmethod
    def torch_uint8():
        return torch.uint8

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            nat.lib.ovl_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _empty(self, n: int, dtype) -> torch.Tensor:
        return torch.empty(max(int(n), 1), dtype=dtype, device=self.device)

    @staticmethod
    def _from_numpy(x: np.ndarray) -> torch.Tensor:
        x = np.ascontiguousarray(x)
        if not x.flags.writeable:          # e.g. np.frombuffer over bytes: torch wants a writable array
            x = x.copy()
        return torch.from_numpy(x)

    def _to_device(self, x, dtype) -> torch.Tensor:
        if isinstance(x, np.ndarray):
            x = self._from_numpy(x)
        if x.dtype != dtype:
            x = x.to(dtype)
        return x.to(self.device, non_blocking=True)

    def _to_device_sharded(self, x, dtype, group, slack: int = 0) -> torch.Tensor:
        """Host array that EVERY rank of `group` holds -> the same array on every rank's GPU, moving it over
        PCIe only once in total: rank r uploads the r-th 1/world of it and one all-gather over NVLink
        completes the copies.  (Each rank uploading everything costs world x the bytes on the host's
        root complexes, which the ranks share.)  Small arrays take the plain path."""
        import torch.distributed as dist
        if isinstance(x, np.ndarray):
            x = self._from_numpy(x)
        world = dist.get_world_size(group) if group is not None else 1
        n = int(x.shape[0])
        if world == 1 or x.is_cuda or x.dtype != dtype or n * x.element_size() < (1 << 20):
            out = torch.empty(n + slack, dtype=dtype, device=self.device)
            out[:n].copy_(x.to(dtype) if x.dtype != dtype else x, non_blocking=True)
            return out[:n] if slack == 0 else out
        rank = dist.get_rank(group)
        gran = max(1, 16 // x.element_size())
        per = ((n + world - 1) // world + gran - 1) // gran * gran
        out = torch.empty(world * per + slack, dtype=dtype, device=self.device)
        lo = min(n, rank * per)
        hi = min(n, lo + per)
        if hi > lo:
            out[lo:hi].copy_(x[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(out[:world * per], out[rank * per:(rank + 1) * per], group=group)
        return out[:n] if slack == 0 else out

    # ------------------------------------------------------------------ K0
    def upload_reads(self, bases, offsets, max_len: Optional[int] = None, code_bits: int = 2) -> ReadSet:
        """bases: uint8[sum len] ASCII, offsets: int64[U+1] (NumPy or CPU/GPU torch tensors)."""
        ascii_dev, off_dev, U, max_len = self._upload_ascii(bases, offsets, max_len)
        return self.pack_reads(ascii_dev, off_dev, U, max_len, code_bits)

    def _upload_ascii(self, bases, offsets, max_len: Optional[int] = None, group=None):
        """Host (or device) reads -> (ascii_dev with slack, off_dev, U, max_len).  With `group` (every rank
        holds the same host arrays) each rank uploads its 1/world and an all-gather completes the copies."""
        off_host = None
        if isinstance(offsets, np.ndarray):
            off_host = offsets
        elif not offsets.is_cuda:
            off_host = offsets.numpy()
        U = int(offsets.shape[0]) - 1
        if U < 0:
            raise ValueError("offsets must have at least one entry")
        if max_len is None:
            if off_host is None:
                off_host = offsets.cpu().numpy()
            max_len = int(np.max(off_host[1:] - off_host[:-1])) if U > 0 else 0
        total = int(bases.shape[0]) if U > 0 else 0
        if group is not None and total:
            src = self._from_numpy(bases) if isinstance(bases, np.ndarray) else bases
            ascii_dev = self._to_device_sharded(src[:total], torch.uint8, group, slack=64)
            off_dev = self._to_device_sharded(offsets, torch.int64, group)
            return ascii_dev, off_dev, U, max_len
        ascii_dev = torch.empty(total + 64, dtype=torch.uint8, device=self.device)   # 32 B slack for K0
        if total:
            src = self._from_numpy(bases) if isinstance(bases, np.ndarray) else bases
            ascii_dev[:total].copy_(src[:total], non_blocking=True)
        off_dev = self._to_device(offsets, torch.int64)
        return ascii_dev, off_dev, U, max_len

    def pack_reads(self, ascii_dev: torch.Tensor, off_dev: torch.Tensor, U: int, max_len: int,
                   code_bits: int = 2) -> ReadSet:
        if max_len > nat.OVL_MAX_LONG_READ_LEN:
            raise nat.OvlUnsupported(f"read length {max_len} exceeds the supported maximum {nat.OVL_MAX_LONG_READ_LEN}")
        if code_bits == 8:
            # any alphabet: padded byte rows, general (slower) kernels downstream
            row_words = max(4, ((max_len + 3) // 4 + 3) // 4 * 4)
            rows = self._empty(U * row_words * 4 + 16, torch.uint8)
            length = self._empty(U, torch.int32)
            bad = torch.zeros(1, dtype=torch.int32, device=self.device)
            nat.check(nat.lib.ovl_pack_bytes(self._ctx, _ptr(ascii_dev), _ptr(off_dev), U, row_words, _ptr(rows),
                                             _ptr(length), self._stream()))
            return ReadSet(rows, length, bad, row_words, U, max_len, 8)
        row_words = int(nat.lib.ovl_row_words(max_len))
        packed = self._empty(U * row_words * 4 + 16, torch.uint8)
        length = self._empty(U, torch.int32)
        bad = torch.zeros(1, dtype=torch.int32, device=self.device)
        nat.check(nat.lib.ovl_pack_reads(self._ctx, _ptr(ascii_dev), _ptr(off_dev), U, row_words,
                                         _ptr(packed), _ptr(length), _ptr(bad), self._stream()))
        return ReadSet(packed, length, bad, row_words, U, max_len)

    def check_alphabet(self, rs: ReadSet) -> None:
        """The reference compares arbitrary characters; the 2-bit kernels cover A/C/G/T only
        and refuse anything else instead of mis-scoring it (host sync)."""
        if rs.code_bits == 2 and int(rs.bad.item()) != 0:
            raise nat.OvlBadAlphabet("reads contain characters other than A, C, G, T; "
                                     "the 2-bit CUDA path does not support them")

    def _total_and_alphabet(self, rs: ReadSet, total_dev: torch.Tensor) -> int:
        """ONE host sync for the two scalars the host needs before it can size the pair list: the pair
        count, and whether the 2-bit packing met a symbol it cannot code (then nothing downstream of
        the mis-coded rows is allocated or run)."""
        both = torch.stack((total_dev.reshape(()).to(torch.int64), rs.bad.reshape(()).to(torch.int64))).cpu()
        if rs.code_bits == 2 and int(both[1]) != 0:
            raise nat.OvlBadAlphabet("reads contain characters other than A, C, G, T; "
                                     "the 2-bit CUDA path does not support them")
        return int(both[0])

    # ------------------------------------------------------------------ K1 + K2
    def kmer_index(self, rs: ReadSet, k: int, segments: Optional[torch.Tensor] = None,
                   n_segments: int = 1) -> KmerIndex:
        """Prefix index (overlapGraphs.py:30-40).  `segments` (int32[U] device tensor) tags every read
        with its read set; reads of different sets then never share a key."""
        if k < 1:
            raise ValueError("k must be positive for the k-mer index")
        hashed = k > nat.OVL_MAX_K or rs.code_bits == 8   # no u64 k-mer: index 64-bit hashes, verify in the join
        key_bits = 64 if hashed else 0
        if segments is not None and hashed:
            raise nat.OvlUnsupported(f"batched read sets with k={k} > {nat.OVL_MAX_K} are not supported")
        if segments is not None:
            seg_bits = max(1, int(n_segments - 1).bit_length())
            key_bits = 2 * k + seg_bits
            if key_bits > 64:
                raise nat.OvlUnsupported(f"k={k} with {n_segments} read sets needs {key_bits} key bits (> 64)")
        U = rs.n_reads
        pk = self._empty(U, torch.int64)
        sk = self._empty(U, torch.int64)
        if rs.code_bits == 8:
            nat.check(nat.lib.ovl_kmer_hashes8(self._ctx, _ptr(rs.packed), rs.row_words, _ptr(rs.length), U, k,
                                               _ptr(pk), _ptr(sk), self._stream()))
        elif hashed:
            nat.check(nat.lib.ovl_kmer_hashes(self._ctx, _ptr(rs.packed), rs.row_words, _ptr(rs.length), U, k,
                                              _ptr(pk), _ptr(sk), self._stream()))
        else:
            nat.check(nat.lib.ovl_kmer_keys(self._ctx, _ptr(rs.packed), rs.row_words, _ptr(rs.length), U, k,
                                            _ptr(segments), _ptr(pk), _ptr(sk), self._stream()))
        sorted_key = self._empty(U, torch.int64)
        sorted_uid = self._empty(U, torch.int32)
        n_indexed = torch.zeros(1, dtype=torch.int64, device=self.device)
        ws_bytes = int(nat.lib.ovl_index_workspace_bytes(U))
        ws = self._empty(ws_bytes, torch.uint8)
        kb = key_bits if key_bits else 2 * k
        table = pos_of = None
        table_bits = 0
        if not hashed:
            # direct-address bucket table + every read's own slot: the join then needs no search
            table_bits = int(nat.lib.ovl_index_table_bits(U, kb))
            table = self._empty((1 << table_bits) + 1, torch.int32)
            pos_of = self._empty(U, torch.int32)
        nat.check(nat.lib.ovl_index_build(self._ctx, _ptr(pk), _ptr(rs.length), U, k, key_bits, _ptr(sorted_key), _ptr(sorted_uid),
                                          _ptr(n_indexed), _ptr(table), table_bits, _ptr(pos_of), None, None, _ptr(ws), ws_bytes,
                                          self._stream()))
        return KmerIndex(k, pk, sk, sorted_key, sorted_uid, n_indexed, kb, table, table_bits, pos_of)

    def _check_fits(self, pairs: int, what: str) -> None:
        """Refuse loudly (instead of running the GPU out of memory) when the pair list, its edge
        offsets and the edge rows of `pairs` candidate pairs cannot fit in free HBM."""
        need = int(pairs) * (4 + 4 + 8 + 16) + (64 << 20)
        if need < self._total_mem // 4:
            return                                 # cudaMemGetInfo costs milliseconds: only ask when it can matter
        free, _ = torch.cuda.mem_get_info(self.device)
        free += torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        if need > free:
            raise nat.OvlUnsupported(f"{what}: {pairs:,} candidate pairs need about {need / 2**30:.1f} GiB of HBM "
                                     f"({free / 2**30:.1f} GiB free); use a larger k or shard over more GPUs")

    # ------------------------------------------------------------------ K3
    def candidate_pairs(self, rs: ReadSet, index: Optional[KmerIndex], k: int,
                        shard: Tuple[int, int] = (0, 1)) -> Tuple[torch.Tensor, torch.Tensor, int]:
        """Ordered candidate list (overlapGraphs.py:43-52).  ``shard=(rank, world)`` returns the
        rank-th contiguous slice of the global (a, b)-ordered list, balanced by pair count.
        Returns (pair_a, pair_b, first_pair_index)."""
        U = rs.n_reads
        rank, world = shard
        st = self._stream()
        if k == 0:
            self.check_alphabet(rs)                            # before U*(U-1) pairs are sized
            total = U * (U - 1) if U > 1 else 0
            p_begin, p_end = total * rank // world, total * (rank + 1) // world
            P = p_end - p_begin
            self._check_fits(P, "k = 0 (all ordered pairs)")
            pair_a = self._empty(P, torch.int32)
            pair_b = self._empty(P, torch.int32)
            if P:
                nat.check(nat.lib.ovl_all_pairs_fill(self._ctx, U, 0, p_begin, P, _ptr(pair_a), _ptr(pair_b), st))
            return pair_a[:P], pair_b[:P], p_begin
        assert index is not None and index.k == k
        if k > nat.OVL_MAX_K or rs.code_bits == 8:
            return self._candidate_pairs_hashed(rs, index, k, shard)
        lo = self._empty(U, torch.int32)
        self_rank = self._empty(U, torch.int32)
        pair_off = self._empty(U + 1, torch.int64)
        ws_bytes = int(nat.lib.ovl_join_workspace_bytes(U))
        ws = self._empty(ws_bytes, torch.uint8)
        nat.check(nat.lib.ovl_join_count(self._ctx, _ptr(index.suffix_key), _ptr(index.prefix_key),
                                         _ptr(rs.length), k, U,
                                         _ptr(index.sorted_key), _ptr(index.sorted_uid), _ptr(index.n_indexed),
                                         _ptr(index.table), index.table_bits, index.key_bits, _ptr(index.pos_of),
                                         None, None, None, _ptr(lo), _ptr(self_rank), _ptr(pair_off), None,
                                         _ptr(ws), ws_bytes, st))
        total = self._total_and_alphabet(rs, pair_off[U])    # host sync: the output size
        p_begin, p_end = total * rank // world, total * (rank + 1) // world
        P = p_end - p_begin
        self._check_fits(P, f"k = {k}")
        pair_a = self._empty(P, torch.int32)
        pair_b = self._empty(P, torch.int32)
        if P:
            nat.check(nat.lib.ovl_join_fill(self._ctx, _ptr(pair_off), 0, U, _ptr(lo), _ptr(self_rank),
                                            _ptr(index.sorted_uid), p_begin, P, total, _ptr(pair_a), _ptr(pair_b), st))
        return pair_a[:P], pair_b[:P], p_begin

    def _candidate_pairs_hashed(self, rs: ReadSet, index: KmerIndex, k: int, shard: Tuple[int, int]):
        """k > OVL_MAX_K: the index holds hashes; every hash match is verified base by base."""
        U = rs.n_reads
        rank, world = shard
        st = self._stream()
        pair_off = self._empty(U + 1, torch.int64)
        ws_bytes = int(nat.lib.ovl_join_workspace_bytes(U))
        ws = self._empty(ws_bytes, torch.uint8)
        count_fn = nat.lib.ovl_join_count_verify8 if rs.code_bits == 8 else nat.lib.ovl_join_count_verify
        fill_fn = nat.lib.ovl_join_fill_verify8 if rs.code_bits == 8 else nat.lib.ovl_join_fill_verify
        nat.check(count_fn(self._ctx, _ptr(rs.packed), rs.row_words, _ptr(rs.length), k,
                                                _ptr(index.suffix_key), 0, U, _ptr(index.sorted_key),
                                                _ptr(index.sorted_uid), _ptr(index.n_indexed), _ptr(pair_off),
                                                _ptr(ws), ws_bytes, st))
        total = self._total_and_alphabet(rs, pair_off[U])
        p_begin, p_end = total * rank // world, total * (rank + 1) // world
        P = p_end - p_begin
        self._check_fits(P, f"k = {k}")
        pair_a = self._empty(P, torch.int32)
        pair_b = self._empty(P, torch.int32)
        if P:
            nat.check(fill_fn(self._ctx, _ptr(rs.packed), rs.row_words, _ptr(rs.length), k,
                                                   _ptr(index.suffix_key), 0, U, _ptr(index.sorted_key),
                                                   _ptr(index.sorted_uid), _ptr(index.n_indexed), _ptr(pair_off),
                                                   p_begin, P, _ptr(pair_a), _ptr(pair_b), st))
        return pair_a[:P], pair_b[:P], p_begin

    # ------------------------------------------------------------------ K0-K3 as one job
    def composite_ok(self, k: int, code_bits: int = 2) -> bool:
        """The one-call K0-K3 job covers 2-bit packed reads with 1 <= k <= OVL_MAX_K."""
        return code_bits == 2 and 1 <= k <= nat.OVL_MAX_K

    def build_candidates(self, ascii_dev: torch.Tensor, off_dev: torch.Tensor, U: int, max_len: int, k: int,
                         copies: Optional[torch.Tensor] = None, node_off: Optional[torch.Tensor] = None,
                         shard: Tuple[int, int] = (0, 1), segments: Optional[torch.Tensor] = None,
                         n_segments: int = 1) -> Candidates:
        """ASCII reads in HBM -> everything up to the SIZED pair list, in ONE library call
        (ovl_candidates_build: pack + keys, index + bucket table, join count, totals) and ONE host
        sync: the totals land in page-locked memory and the host waits on an event.  Raises
        OvlBadAlphabet before anything is sized on mis-coded rows."""
        lay = nat.CandLayout()
        nat.check(nat.lib.ovl_candidates_layout(U, max_len, k, int(n_segments) if segments is not None else 1,
                                                1 if copies is not None else 0, ctypes.byref(lay)))
        arena = torch.empty(int(lay.total_bytes), dtype=torch.uint8, device=self.device)
        rank, world = shard
        nat.check(nat.lib.ovl_candidates_build(self._ctx, _ptr(ascii_dev), _ptr(off_dev), U, k, _ptr(segments), _ptr(copies),
                                               int(rank), int(world), _ptr(arena), ctypes.byref(lay),
                                               ctypes.c_void_p(self._totals.data_ptr()), self._stream()))
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.device))
        n = max(U, 1)

        def view(off, count, dtype):
            nbytes = count * torch.empty(0, dtype=dtype).element_size()
            return arena[off:off + nbytes].view(dtype)

        rs = ReadSet(arena[lay.packed:lay.packed + n * lay.row_words * 4 + 16], view(lay.len, n, torch.int32),
                     view(lay.bad, 1, torch.int32), int(lay.row_words), U, max_len)
        index = KmerIndex(k, view(lay.prefix_key, n, torch.int64), view(lay.suffix_key, n, torch.int64),
                          view(lay.sorted_key, n, torch.int64), view(lay.sorted_uid, n, torch.int32),
                          view(lay.n_indexed, 1, torch.int64), int(lay.key_bits),
                          view(lay.table, (1 << lay.table_bits) + 1, torch.int32), int(lay.table_bits), None)
        has = copies is not None
        done.synchronize()                                            # the one host sync of the job
        t = self._totals.tolist()
        if t[2] != 0:
            raise nat.OvlBadAlphabet("reads contain characters other than A, C, G, T; "
                                     "the 2-bit CUDA path does not support them")
        nc = 65
        return Candidates(rs, index, view(lay.bucket_lo, n, torch.int32), view(lay.self_rank, n, torch.int32),
                          view(lay.pair_off, n + 1, torch.int64),
                          view(lay.edge_base, n + 1, torch.int64) if has else None,
                          view(lay.cum, n + 1, torch.int64) if has else None,
                          copies, node_off, int(t[0]), int(t[1]), int(t[3]), int(t[4]),
                          [int(x) for x in t[8:8 + nc]], [int(x) for x in t[8 + nc:8 + 2 * nc]], arena)

    def fill_pairs(self, cand: Candidates) -> Tuple[torch.Tensor, torch.Tensor]:
        """This rank's slice of the ordered candidate list (overlapGraphs.py:43-52)."""
        P = cand.p_end - cand.p_begin
        self._check_fits(P, f"k = {cand.index.k}")
        pair_a = self._empty(P, torch.int32)
        pair_b = self._empty(P, torch.int32)
        if P:
            nat.check(nat.lib.ovl_join_fill(self._ctx, _ptr(cand.pair_off), 0, cand.rs.n_reads, _ptr(cand.bucket_lo),
                                            _ptr(cand.self_rank), _ptr(cand.index.sorted_uid), cand.p_begin, P,
                                            cand.total_pairs, _ptr(pair_a), _ptr(pair_b), self._stream()))
        return pair_a[:P], pair_b[:P]

    def _dp_edges_join_call(self, cand: Candidates, a_ptr, b_ptr, P: int, p_begin: int, e_begin: int, scoring, out_ptr, st):
        """DP + fused edge expansion with the row offsets taken from the join index (no per-pair offset array)."""
        match_score, mismatch, indel = scoring
        rs = cand.rs
        nat.check(nat.lib.ovl_overlap_dp_edges_join(self._ctx, _ptr(rs.packed), rs.row_words, _ptr(rs.length), a_ptr, b_ptr, P,
                                                    rs.max_len, int(match_score), int(mismatch), int(indel),
                                                    _ptr(cand.copies), _ptr(cand.node_off), _ptr(cand.pair_off),
                                                    _ptr(cand.edge_base), _ptr(cand.bucket_lo), _ptr(cand.self_rank),
                                                    _ptr(cand.cum), int(p_begin), int(e_begin), out_ptr, st))

    def candidate_edges(self, cand: Candidates, pair_a: torch.Tensor, pair_b: torch.Tensor,
                        match_score: int = 10, mismatch: int = -1, indel: int = INDEL_DEFAULT,
                        events=None, sink=None) -> Optional[torch.Tensor]:
        """DP + edge expansion (overlapGraphs.py:53-60) over this rank's slice of a build_candidates job:
        device edge rows, or -- with `sink` -- rows stored wherever sink(E) points (peer memory)."""
        P = int(pair_a.shape[0])
        E = cand.e_end - cand.e_begin
        st = self._stream()
        if P == 0:
            if sink is not None:
                sink(0, cand.e_begin, cand.total_edges)
            return torch.empty((0, 4), dtype=torch.int32, device=self.device)
        if sink is not None:
            # every rank knows its global row offset from the join index: no size exchange
            edges = None
            out_ptr = ctypes.c_void_p(int(sink(E, cand.e_begin, cand.total_edges)))
        else:
            edges = self._empty(E * 4, torch.int32)
            out_ptr = _ptr(edges)
        if events is not None:
            events[0].record()
        if cand.edge_base is not None:
            self._dp_edges_join_call(cand, _ptr(pair_a), _ptr(pair_b), P, cand.p_begin, cand.e_begin,
                                     (match_score, mismatch, indel), out_ptr, st)
        else:
            self._dp_edges_call(cand.rs, _ptr(pair_a), _ptr(pair_b), P, (match_score, mismatch, indel), None, None,
                                ctypes.c_void_p(0), out_ptr, st)
        if events is not None:
            events[1].record()
        return None if edges is None else edges[:E * 4].view(E, 4)

    def candidate_edges_to_host(self, cand: Candidates, pair_a: torch.Tensor, pair_b: torch.Tensor,
                                match_score: int = 10, mismatch: int = -1, indel: int = INDEL_DEFAULT,
                                chunk_pairs: int = 1 << 62, host_sink=None, min_chunked_pairs: int = 1_000_000) -> np.ndarray:
        """The same with the device->host copy of the edge rows overlapped with the DP: the slice is cut at
        boundaries the job already computed (cand.cut_pairs / cut_edges), each chunk's rows are copied
        to the pinned host buffer on a second stream while the next chunk computes."""
        P = int(pair_a.shape[0])
        if P == 0:
            if host_sink is not None:
                host_sink(0, cand.e_begin, cand.total_edges)
            return np.zeros((0, 4), np.int32)
        main = torch.cuda.current_stream(self.device)
        st = self._stream()
        cuts = d2h_chunk_cuts(P, chunk_pairs, min_chunked_pairs)
        n_chunks = len(cuts) - 1
        bounds = [cand.cut_pairs[c] - cand.p_begin for c in cuts]
        e_bounds = [cand.cut_edges[c] - cand.e_begin for c in cuts]
        E = e_bounds[-1]
        edges = self._empty(E * 4, torch.int32).view(-1, 4)
        host = host_sink(E, cand.e_begin, cand.total_edges) if host_sink is not None else self._host_rows(E)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(self.device)
        scoring = (match_score, mismatch, indel)
        for c in range(n_chunks):
            p0, p1 = bounds[c], bounds[c + 1]
            if p1 == p0:
                continue
            e0, e1 = e_bounds[c], e_bounds[c + 1]
            a_ptr = ctypes.c_void_p(pair_a.data_ptr() + 4 * p0)
            b_ptr = ctypes.c_void_p(pair_b.data_ptr() + 4 * p0)
            out_ptr = ctypes.c_void_p(edges.data_ptr() + 16 * e0)
            if cand.edge_base is not None:
                self._dp_edges_join_call(cand, a_ptr, b_ptr, p1 - p0, cand.p_begin + p0, cand.e_begin + e0, scoring, out_ptr, st)
            else:
                self._dp_edges_call(cand.rs, a_ptr, b_ptr, p1 - p0, scoring, None, None, ctypes.c_void_p(0), out_ptr, st)
            done = torch.cuda.Event()
            done.record(main)
            if e1 > e0:
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(done)
                    host[e0:e1].copy_(edges[e0:e1], non_blocking=True)
        self._copy_stream.synchronize()
        main.wait_stream(self._copy_stream)
        edges.record_stream(self._copy_stream)
        return host.numpy()

    def _host_rows(self, E: int, host_sink=None) -> torch.Tensor:
        """Page-locked destination for E edge rows: the caller's sink or the engine's reusable buffer."""
        if host_sink is not None:
            return host_sink(E)
        if self._pinned_out is None or self._pinned_out.shape[0] < max(E, 1):
            self._pinned_out = torch.empty((max(E, 1) * 5 // 4 + 16, 4), dtype=torch.int32).pin_memory()
        return self._pinned_out[:E]

    # ------------------------------------------------------------------ K4 / K5
    def overlap_scores(self, rs: ReadSet, pair_a: torch.Tensor, pair_b: torch.Tensor,
                       match_score: int = 10, mismatch: int = -1, indel: int = INDEL_DEFAULT,
                       mode: int = 0, lanes: int = 0, cols: int = 0,
                       out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(score[p], end[p]) of overlap_alignment(read[a[p]], read[b[p]]) -- overlapGraphs.py:53."""
        P = int(pair_a.shape[0])
        if out is None:
            score = self._empty(P, torch.int32)
            end = self._empty(P, torch.int32)
        else:
            score, end = out
        if P and rs.code_bits == 8:
            nat.check(nat.lib.ovl_overlap_dp8(self._ctx, _ptr(rs.packed), rs.row_words, _ptr(rs.length),
                                              _ptr(pair_a), _ptr(pair_b), P, rs.max_len,
                                              int(match_score), int(mismatch), int(indel), _ptr(score), _ptr(end),
                                              None, None, None, None, self._stream()))
        elif P:
            nat.check(nat.lib.ovl_overlap_dp(self._ctx, _ptr(rs.packed), rs.row_words, _ptr(rs.length),
                                             _ptr(pair_a), _ptr(pair_b), P, rs.max_len,
                                             int(match_score), int(mismatch), int(indel),
                                             _ptr(score), _ptr(end), mode, lanes, cols, self._stream()))
        return score[:P], end[:P]

    def _dp_edges_call(self, rs: ReadSet, a_ptr, b_ptr, P: int, scoring, copies, node_off, off_ptr, out_ptr, st) -> None:
        """One launch of the DP with the fused edge epilogue (2-bit packed or byte-coded reads)."""
        match_score, mismatch, indel = scoring
        if rs.code_bits == 8:
            nat.check(nat.lib.ovl_overlap_dp8(self._ctx, _ptr(rs.packed), rs.row_words, _ptr(rs.length), a_ptr, b_ptr, P,
                                              rs.max_len, int(match_score), int(mismatch), int(indel), None, None,
                                              _ptr(copies), _ptr(node_off), off_ptr, out_ptr, st))
        else:
            nat.check(nat.lib.ovl_overlap_dp_edges(self._ctx, _ptr(rs.packed), rs.row_words, _ptr(rs.length), a_ptr, b_ptr,
                                                   P, rs.max_len, int(match_score), int(mismatch), int(indel),
                                                   _ptr(copies), _ptr(node_off), off_ptr, out_ptr, st))

    def overlap_edges_fused(self, rs: ReadSet, pair_a: torch.Tensor, pair_b: torch.Tensor,
                            copies: Optional[torch.Tensor] = None, node_off: Optional[torch.Tensor] = None,
                            match_score: int = 10, mismatch: int = -1, indel: int = INDEL_DEFAULT,
                            events=None, sink=None) -> Optional[torch.Tensor]:
        """DP + edge expansion in one kernel (overlapGraphs.py:53-60): the DP epilogue writes the
        copy_a x copy_b edge rows, so score/end never travel through HBM."""
        P = int(pair_a.shape[0])
        st = self._stream()
        if P == 0:
            if sink is not None:
                sink(0)                            # collective sinks must be called by every rank
            return torch.empty((0, 4), dtype=torch.int32, device=self.device)
        edge_off = None
        E = P
        if copies is not None:
            edge_off = self._empty(P + 1, torch.int64)
            ws_bytes = int(nat.lib.ovl_expand_workspace_bytes(P))
            ws = self._empty(ws_bytes, torch.uint8)
            nat.check(nat.lib.ovl_expand_count(self._ctx, _ptr(pair_a), _ptr(pair_b), _ptr(copies), P,
                                               _ptr(edge_off), _ptr(ws), ws_bytes, st))
            E = int(edge_off[P].item())                       # host sync: the output size
        if sink is not None:
            # `sink(E)` returns a raw device address -- possibly in a PEER GPU's memory (NVLink): the
            # kernel's epilogue stores the rows there directly (parallel.PeerEdgeBuffer)
            edges = None
            out_ptr = ctypes.c_void_p(int(sink(E)))
        else:
            edges = self._empty(E * 4, torch.int32)
            out_ptr = _ptr(edges)
        if events is not None:
            events[0].record()
        self._dp_edges_call(rs, _ptr(pair_a), _ptr(pair_b), P, (match_score, mismatch, indel), copies, node_off,
                            _ptr(edge_off), out_ptr, st)
        if events is not None:
            events[1].record()
        return None if edges is None else edges[:E * 4].view(E, 4)

    def overlap_edges_fused_to_host(self, rs: ReadSet, pair_a: torch.Tensor, pair_b: torch.Tensor,
                                    copies: Optional[torch.Tensor] = None, node_off: Optional[torch.Tensor] = None,
                                    match_score: int = 10, mismatch: int = -1, indel: int = INDEL_DEFAULT,
                                    chunk_pairs: int = 48_000_000, host_sink=None) -> np.ndarray:
        """DP + fused edge expansion, with the device->host copy of the edge rows overlapped with the
        DP: the pair list is cut into chunks, each chunk's rows are copied to the pinned host buffer
        on a second stream while the next chunk computes.  Returns a view of the pinned buffer."""
        P = int(pair_a.shape[0])
        if P == 0:
            if host_sink is not None:
                host_sink(0)                       # collective sinks must be called by every rank
            return np.zeros((0, 4), np.int32)
        main = torch.cuda.current_stream(self.device)
        st = self._stream()
        n_chunks = max(1, min(64, (P + chunk_pairs - 1) // chunk_pairs))
        if P >= 1_000_000:
            n_chunks = max(n_chunks, 8)        # even a few-ms job hides most of its copy behind the DP
        bounds = [P * c // n_chunks for c in range(n_chunks + 1)]
        edge_off = None
        if copies is not None:
            edge_off = self._empty(P + 1, torch.int64)
            ws_bytes = int(nat.lib.ovl_expand_workspace_bytes(P))
            ws = self._empty(ws_bytes, torch.uint8)
            nat.check(nat.lib.ovl_expand_count(self._ctx, _ptr(pair_a), _ptr(pair_b), _ptr(copies), P,
                                               _ptr(edge_off), _ptr(ws), ws_bytes, st))
            idx = torch.tensor(bounds, dtype=torch.int64, device=self.device)
            e_bounds = edge_off[idx].cpu().tolist()           # host sync: output size + chunk boundaries
        else:
            e_bounds = bounds
        E = int(e_bounds[-1])
        edges = self._empty(E * 4, torch.int32).view(-1, 4)
        if host_sink is not None:
            host = host_sink(E)                    # caller-provided page-locked destination (E rows)
        else:
            if self._pinned_out is None or self._pinned_out.shape[0] < max(E, 1):
                self._pinned_out = torch.empty((max(E, 1) * 5 // 4 + 16, 4), dtype=torch.int32).pin_memory()
            host = self._pinned_out[:E]
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(self.device)
        for c in range(n_chunks):
            p0, p1 = bounds[c], bounds[c + 1]
            if p1 == p0:
                continue
            a_ptr = ctypes.c_void_p(pair_a.data_ptr() + 4 * p0)
            b_ptr = ctypes.c_void_p(pair_b.data_ptr() + 4 * p0)
            if copies is not None:
                off_ptr = ctypes.c_void_p(edge_off.data_ptr() + 8 * p0)       # offsets stay global
                out_ptr = _ptr(edges)
            else:
                off_ptr = ctypes.c_void_p(0)
                out_ptr = ctypes.c_void_p(edges.data_ptr() + 16 * p0)
            self._dp_edges_call(rs, a_ptr, b_ptr, p1 - p0, (match_score, mismatch, indel), copies, node_off,
                                off_ptr, out_ptr, st)
            done = torch.cuda.Event()
            done.record(main)
            e0, e1 = int(e_bounds[c]), int(e_bounds[c + 1])
            if e1 > e0:
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(done)
                    host[e0:e1].copy_(edges[e0:e1], non_blocking=True)
        self._copy_stream.synchronize()
        main.wait_stream(self._copy_stream)
        edges.record_stream(self._copy_stream)
        return host.numpy()

    @staticmethod
    def dp_plan(max_len: int, match_score: int = 10, mismatch: int = -1, indel: int = INDEL_DEFAULT, mode: int = 0):
        out = (ctypes.c_int32 * 3)()
        nat.check(nat.lib.ovl_overlap_dp_plan(max_len, int(match_score), int(mismatch), int(indel), mode,
                                              ctypes.byref(out)))
        return {"mode": {1: "packed16", 2: "int32", 3: "long-read"}[int(out[0])], "lanes": int(out[1]), "cols": int(out[2])}

    # ------------------------------------------------------------------ K7
    def align_pair(self, s_codes: np.ndarray, t_codes: np.ndarray, match_score: int = 10, mismatch: int = -1,
                   indel: int = INDEL_DEFAULT) -> Tuple[int, int, np.ndarray]:
        """One pair with traceback (aligners.py:27-76).  s_codes / t_codes are int32 code
        points.  Returns (best_score, end_position, ops) with ops the traceback from the end
        backwards (0 diagonal, 1 up, 2 left)."""
        n, m = int(s_codes.shape[0]), int(t_codes.shape[0])
        both = np.concatenate([s_codes.astype(np.int32), t_codes.astype(np.int32), np.zeros(1, np.int32)])
        dev = self._to_device(both, torch.int32)
        ws_bytes = int(nat.lib.ovl_align_pair_workspace_bytes(n, m))
        ws = self._empty(ws_bytes, torch.uint8)
        result = torch.zeros(4, dtype=torch.int32, device=self.device)
        ops = self._empty(n + m + 1, torch.uint8)
        s_ptr = ctypes.c_void_p(dev.data_ptr())
        t_ptr = ctypes.c_void_p(dev.data_ptr() + 4 * n)
        nat.check(nat.lib.ovl_align_pair(self._ctx, s_ptr, n, t_ptr, m, int(match_score), int(mismatch), int(indel),
                                         _ptr(ws), ws_bytes, _ptr(result), _ptr(ops), self._stream()))
        res = result.cpu().numpy()
        n_ops = int(res[2])
        return int(res[0]), int(res[1]), ops[:n_ops].cpu().numpy()

    # ------------------------------------------------------------------ K8
    def local_align(self, q_codes: np.ndarray, r_codes: np.ndarray, match_score: int = 10, mismatch: int = -1,
                    indel: int = -1):
        """Smith-Waterman with traceback (aligners.py:85-167).  Returns (best_score, start_pos, end_pos,
        best_i, ops) with ops from the best cell backwards (1 diagonal, 2 up, 3 left)."""
        n, m = int(q_codes.shape[0]), int(r_codes.shape[0])
        both = np.concatenate([q_codes.astype(np.int32), r_codes.astype(np.int32), np.zeros(1, np.int32)])
        dev = self._to_device(both, torch.int32)
        ws_bytes = int(nat.lib.ovl_local_align_workspace_bytes(n, m))
        ws = self._empty(ws_bytes, torch.uint8)
        result = torch.zeros(8, dtype=torch.int32, device=self.device)
        ops = self._empty(n + m + 1, torch.uint8)
        nat.check(nat.lib.ovl_local_align(self._ctx, ctypes.c_void_p(dev.data_ptr()), n,
                                          ctypes.c_void_p(dev.data_ptr() + 4 * n), m,
                                          int(match_score), int(mismatch), int(indel),
                                          _ptr(ws), ws_bytes, _ptr(result), _ptr(ops), self._stream()))
        res = result.cpu().numpy()
        return int(res[0]), int(res[1]), int(res[2]), int(res[4]), ops[:int(res[3])].cpu().numpy()

    def local_align_batch(self, queries, r_codes: np.ndarray, windows, match_score: int = 10, mismatch: int = -1,
                          indel: int = -1):
        """Many queries (int32 code arrays, each <= OVL_LOCAL_BATCH_MAX_QUERY symbols) against windows
        (start, length) of one reference, ONE launch with one CTA per query (aligners.py:85-167 each).
        Returns [(best_score, start_pos, end_pos, best_i, ops)] with positions relative to the window."""
        nq = len(queries)
        if nq == 0:
            return []
        lens = np.fromiter((len(q) for q in queries), dtype=np.int64, count=nq)
        if int(lens.max()) > nat.OVL_LOCAL_BATCH_MAX_QUERY:
            raise nat.OvlUnsupported(f"batched local alignment takes queries of at most {nat.OVL_LOCAL_BATCH_MAX_QUERY} symbols")
        q_off = np.zeros(nq + 1, np.int64)
        np.cumsum(lens, out=q_off[1:])
        w = np.asarray(windows, dtype=np.int64).reshape(nq, 2)
        m = w[:, 1]
        tb_sz = ((lens + 1) * (lens + m + 2) + 15) // 16 * 16
        ops_sz = (lens + m + 1 + 15) // 16 * 16
        tb_off = np.zeros(nq + 1, np.int64)
        np.cumsum(tb_sz, out=tb_off[1:])
        ops_off = np.zeros(nq + 1, np.int64)
        np.cumsum(ops_sz, out=ops_off[1:])
        q_all = np.concatenate([np.asarray(q, dtype=np.int32) for q in queries] + [np.zeros(1, np.int32)])
        d_q = self._to_device(q_all, torch.int32)
        d_ref = self._to_device(np.concatenate([r_codes.astype(np.int32), np.zeros(1, np.int32)]), torch.int32)
        meta = np.concatenate([q_off, tb_off[:-1], ops_off[:-1]])
        d_meta = self._to_device(meta, torch.int64)
        d_win = self._to_device(np.concatenate([w[:, 0], w[:, 1]]).astype(np.int32), torch.int32)
        tb = self._empty(int(tb_off[-1]), torch.uint8)
        ops = self._empty(int(ops_off[-1]), torch.uint8)
        results = torch.zeros(nq * 8, dtype=torch.int32, device=self.device)
        base = d_meta.data_ptr()
        nat.check(nat.lib.ovl_local_align_batch(
            self._ctx, _ptr(d_q), ctypes.c_void_p(base), nq, int(lens.max()), _ptr(d_ref),
            ctypes.c_void_p(d_win.data_ptr()), ctypes.c_void_p(d_win.data_ptr() + 4 * nq),
            int(match_score), int(mismatch), int(indel), _ptr(tb), ctypes.c_void_p(base + 8 * (nq + 1)),
            _ptr(results), _ptr(ops), ctypes.c_void_p(base + 8 * (2 * nq + 1)), self._stream()))
        res = results.cpu().numpy().reshape(nq, 8)
        ops_h = ops.cpu().numpy()
        out = []
        for x in range(nq):
            o0 = int(ops_off[x])
            out.append((int(res[x, 0]), int(res[x, 1]), int(res[x, 2]), int(res[x, 4]), ops_h[o0:o0 + int(res[x, 3])]))
        return out

    # ------------------------------------------------------------------ K6
    def expand_edges(self, pair_a, pair_b, score, end, copies: Optional[torch.Tensor] = None,
                     node_off: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Edge rows int32[E, 4] = (node_a, node_b, weight, end_position), overlapGraphs.py:55-60.
        copies=None means every read occurs once (node id == uid)."""
        P = int(pair_a.shape[0])
        st = self._stream()
        if P == 0:
            return torch.empty((0, 4), dtype=torch.int32, device=self.device)
        if copies is None:
            edges = self._empty(P * 4, torch.int32)
            if P:
                nat.check(nat.lib.ovl_expand_unit(self._ctx, _ptr(pair_a), _ptr(pair_b), _ptr(score), _ptr(end),
                                                  P, _ptr(edges), st))
            return edges[:P * 4].view(P, 4)
        edge_off = self._empty(P + 1, torch.int64)
        ws_bytes = int(nat.lib.ovl_expand_workspace_bytes(P))
        ws = self._empty(ws_bytes, torch.uint8)
        nat.check(nat.lib.ovl_expand_count(self._ctx, _ptr(pair_a), _ptr(pair_b), _ptr(copies), P,
                                           _ptr(edge_off), _ptr(ws), ws_bytes, st))
        E = int(edge_off[P].item())                           # host sync: the output size
        edges = self._empty(E * 4, torch.int32)
        if E:
            nat.check(nat.lib.ovl_expand_fill(self._ctx, _ptr(edge_off), P, _ptr(pair_a), _ptr(pair_b),
                                              _ptr(score), _ptr(end), _ptr(copies), _ptr(node_off), 0, E,
                                              _ptr(edges), st))
        return edges[:E * 4].view(E, 4)

    def filter_edges(self, edges: torch.Tensor, min_weight: int) -> torch.Tensor:
        """Keep the edge rows with weight >= min_weight, order preserved (overlapGraphs.py:225, :347)."""
        E = int(edges.shape[0])
        if E == 0:
            return edges
        st = self._stream()
        keep_off = self._empty(E + 1, torch.int64)
        ws_bytes = int(nat.lib.ovl_filter_workspace_bytes(E))
        ws = self._empty(ws_bytes, torch.uint8)
        nat.check(nat.lib.ovl_filter_count(self._ctx, _ptr(edges), E, int(min_weight), _ptr(keep_off), _ptr(ws), ws_bytes, st))
        kept = int(keep_off[E].item())
        out = self._empty(kept * 4, torch.int32)
        nat.check(nat.lib.ovl_filter_fill(self._ctx, _ptr(edges), _ptr(keep_off), E, int(min_weight), _ptr(out), st))
        return out[:kept * 4].view(kept, 4)

    # ------------------------------------------------------------------ cycle-removal pre-pass
    def trim_sinks(self, src, dst, n_nodes: int):
        """Peel sinks off a directed graph (edge list src[e] -> dst[e]) until none is left (csrc/trim.cuh).
        Returns (keep bool[n_nodes], rounds): keep[v] is False for the nodes that cannot reach a cycle."""
        E = int(len(src))
        d_src = self._to_device(np.ascontiguousarray(src, dtype=np.int32), torch.int32) if E else None
        d_dst = self._to_device(np.ascontiguousarray(dst, dtype=np.int32), torch.int32) if E else None
        state = torch.zeros(max(n_nodes, 1), dtype=torch.int32, device=self.device)
        ws_bytes = int(nat.lib.ovl_trim_workspace_bytes(n_nodes))
        ws = self._empty(ws_bytes, torch.uint8)
        rounds = ctypes.c_int32(0)
        nat.check(nat.lib.ovl_trim_sinks(self._ctx, _ptr(d_src), _ptr(d_dst), E, int(n_nodes), _ptr(state), _ptr(ws), ws_bytes,
                                         ctypes.byref(rounds), self._stream()))
        return (state[:n_nodes] == 0).cpu().numpy(), int(rounds.value)

    # ------------------------------------------------------------------ read simulator
    def simulate_reads(self, genome, n_reads: int, read_len: int, error_prob: float, seed: int):
        """Seeded read simulation ON THE DEVICE (generateErrorFreeReads.py:22-52 + generateErrorProneReads.py:4-45
        as a counter-based stream): returns (ascii_dev with 64 bytes of slack, offsets_dev int64[n_reads+1]) --
        directly what build_candidates() takes -- bit-identical to synth.simulate_reads_counter()."""
        from .synth import error_threshold
        g = self._to_device(genome, torch.uint8) if not (isinstance(genome, torch.Tensor) and genome.is_cuda) else genome
        G = int(g.shape[0])
        ascii_dev = torch.empty(n_reads * read_len + 64, dtype=torch.uint8, device=self.device)
        offsets = torch.empty(n_reads + 1, dtype=torch.int64, device=self.device)
        ws_bytes = int(nat.lib.ovl_simulate_workspace_bytes(n_reads))
        ws = self._empty(ws_bytes, torch.uint8)
        nat.check(nat.lib.ovl_simulate_reads(self._ctx, _ptr(g), G, n_reads, read_len, error_threshold(error_prob),
                                             int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(offsets), _ptr(ascii_dev), _ptr(ws),
                                             ws_bytes, self._stream()))
        return ascii_dev, offsets

    # ------------------------------------------------------------------ edge-list fingerprint
    def edge_hash(self, edges: torch.Tensor, first_row: int = 0, accum: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Order-sensitive fingerprint of device edge rows (csrc/check.cuh): returns the int64[1] device
        accumulator (bit pattern of the u64 sum); shards hashed with their global `first_row` add up."""
        if accum is None:
            accum = torch.zeros(1, dtype=torch.int64, device=self.device)
        E = int(edges.shape[0])
        if E:
            nat.check(nat.lib.ovl_edge_list_hash(self._ctx, _ptr(edges), E, int(first_row), _ptr(accum), self._stream()))
        return accum

    def edge_hash_host(self, rows: np.ndarray, first_row: int = 0, chunk_rows: int = 1 << 25) -> int:
        """The same fingerprint of HOST edge rows: uploaded chunk by chunk and hashed by the same kernel."""
        accum = torch.zeros(1, dtype=torch.int64, device=self.device)
        E = int(rows.shape[0])
        for r0 in range(0, E, chunk_rows):
            part = self._from_numpy(rows[r0:r0 + chunk_rows]).to(self.device, non_blocking=True)
            self.edge_hash(part, first_row + r0, accum)
        return int(accum.item()) & 0xFFFFFFFFFFFFFFFF

    @staticmethod
    def edge_hash_numpy(rows: np.ndarray, first_row: int = 0) -> int:
        """Host mirror of csrc/check.cuh (wrapping u64 arithmetic), for tests and small lists."""
        r = np.ascontiguousarray(rows, dtype=np.int32).view(np.uint32).astype(np.uint64)
        n = r.shape[0]
        with np.errstate(over="ignore"):
            z = (np.arange(n, dtype=np.uint64) + np.uint64(first_row)) * np.uint64(0x9E3779B97F4A7C15) \
                + np.uint64(0x632BE59BD9B4E019)
            z ^= r[:, 0] | (r[:, 1] << np.uint64(32))
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z ^= r[:, 2] | (r[:, 3] << np.uint64(32))
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z ^= z >> np.uint64(31)
            return int(z.sum(dtype=np.uint64))

    # ------------------------------------------------------------------ whole path
    def overlap_edges_device(self, rs: ReadSet, k: int, copies: Optional[torch.Tensor] = None,
                             node_off: Optional[torch.Tensor] = None, shard: Tuple[int, int] = (0, 1),
                             match_score: int = 10, mismatch: int = -1, indel: int = INDEL_DEFAULT,
                             stats: Optional[dict] = None, segments: Optional[torch.Tensor] = None,
                             n_segments: int = 1) -> torch.Tensor:
        """Reads already packed in HBM -> device edge rows (this rank's shard)."""
        if segments is not None and k == 0:
            raise nat.OvlUnsupported("batched read sets need k > 0 (k = 0 pairs every read with every other)")
        index = self.kmer_index(rs, k, segments, n_segments) if k > 0 else None
        pair_a, pair_b, _ = self.candidate_pairs(rs, index, k, shard)
        edges = self.overlap_edges_fused(rs, pair_a, pair_b, copies, node_off, match_score, mismatch, indel)
        if stats is not None:
            stats["pairs"] = int(pair_a.shape[0])
            stats["edges"] = int(edges.shape[0])
            stats["pair_a"], stats["pair_b"] = pair_a, pair_b
        return edges

    def to_pinned_host(self, edges: torch.Tensor) -> np.ndarray:
        """Device edge rows -> NumPy view of the engine's reusable pinned buffer (valid until the
        next call)."""
        E = int(edges.shape[0])
        if self._pinned_out is None or self._pinned_out.shape[0] < max(E, 1):
            self._pinned_out = torch.empty((max(E, 1) * 5 // 4 + 16, 4), dtype=torch.int32).pin_memory()
        host = self._pinned_out[:E]
        host.copy_(edges, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host.numpy()

    def overlap_edges(self, bases, offsets, counts=None, k: int = 5, shard: Tuple[int, int] = (0, 1),
                      match_score: int = 10, mismatch: int = -1, indel: int = INDEL_DEFAULT,
                      stats: Optional[dict] = None, to_host: bool = True, reuse_host_buffer: bool = False,
                      min_weight: Optional[int] = None, pairs=None, segments=None, n_segments: int = 1,
                      host_sink=None, code_bits: int = 2, upload_group=None):
        """HOST buffers in, HOST edge rows out: unique reads (ASCII bytes + offsets) and their
        multiplicities -> int32[E, 4] (node_a, node_b, weight, end_position) in the reference's
        insertion order.  This is the call the drop-in graph builder makes.
        upload_group: a torch.distributed group whose ranks all make this call with the SAME host arrays
        (the sharded builder): the inputs then cross PCIe once in total instead of once per rank."""
        if k < 0:
            raise AssertionError("k-mer length must be non-negative")      # overlapGraphs.py:17
        copies = node_off = None
        if counts is not None:
            counts_np = counts if isinstance(counts, np.ndarray) else counts.numpy()
            if counts_np.size and int(counts_np.max()) > 1:
                no = np.zeros(counts_np.shape[0] + 1, dtype=np.int64)
                np.cumsum(counts_np, out=no[1:])
                if upload_group is not None and counts_np.dtype == np.int32:
                    copies = self._to_device_sharded(counts_np, torch.int32, upload_group)
                    node_off = self._to_device_sharded(no, torch.int64, upload_group)
                else:
                    copies = self._to_device(counts_np, torch.int32)
                    node_off = self._to_device(no, torch.int64)
        if pairs is None and self.composite_ok(k, code_bits):
            # the common call: K0-K3 as one job, one host sync, then fill + DP (+ D2H overlapped with the DP)
            ascii_dev, off_dev, U, max_len = self._upload_ascii(bases, offsets, group=upload_group)
            seg_dev = self._to_device(segments, torch.int32) if segments is not None else None
            cand = self.build_candidates(ascii_dev, off_dev, U, max_len, k, copies, node_off, shard, seg_dev, n_segments)
            pa, pb = self.fill_pairs(cand)
            if stats is not None:
                stats["pairs"], stats["edges"] = int(pa.shape[0]), cand.e_end - cand.e_begin
                stats["pair_a"], stats["pair_b"] = pa, pb
            if to_host and min_weight is None:
                host = self.candidate_edges_to_host(cand, pa, pb, match_score, mismatch, indel, host_sink=host_sink)
                return host if (reuse_host_buffer or host_sink is not None) else host.copy()
            edges = self.candidate_edges(cand, pa, pb, match_score, mismatch, indel)
            if min_weight is not None:
                edges = self.filter_edges(edges, min_weight)
            if not to_host:
                return edges
            host = self.to_pinned_host(edges)
            return host if reuse_host_buffer else host.copy()
        rs = self.upload_reads(bases, offsets, code_bits=code_bits)
        if pairs is not None:        # caller-supplied unique-read index pairs instead of the k-mer join
            self.check_alphabet(rs)
            pa = self._to_device(pairs[0], torch.int32)
            pb = self._to_device(pairs[1], torch.int32)
            edges = self.overlap_edges_fused(rs, pa, pb, copies, node_off, match_score, mismatch, indel)
        elif to_host and min_weight is None and segments is None:
            # the common host call: overlap the D2H of the edge rows with the DP, chunk by chunk
            index = self.kmer_index(rs, k) if k > 0 else None
            pa, pb, _ = self.candidate_pairs(rs, index, k, shard)
            host = self.overlap_edges_fused_to_host(rs, pa, pb, copies, node_off, match_score, mismatch, indel,
                                                    host_sink=host_sink)
            if stats is not None:
                stats["pairs"], stats["edges"] = int(pa.shape[0]), int(host.shape[0])
            return host if (reuse_host_buffer or host_sink is not None) else host.copy()
        else:
            seg_dev = self._to_device(segments, torch.int32) if segments is not None else None
            edges = self.overlap_edges_device(rs, k, copies, node_off, shard, match_score, mismatch, indel, stats,
                                              seg_dev, n_segments)
        if min_weight is not None:
            edges = self.filter_edges(edges, min_weight)
        if not to_host:
            return edges
        host = self.to_pinned_host(edges)
        # reuse_host_buffer: hand out the engine's pinned buffer itself (valid until the next call)
        return host if reuse_host_buffer else host.copy()


_ENGINES = {}


def get_engine(device: Optional[int] = None) -> OverlapEngine:
    """Process-wide engine per device (created on first use)."""
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device visible: the overlap path runs on sm_100a only (no CPU fallback)")
    if device is None:
        device = torch.cuda.current_device()
    eng = _ENGINES.get(device)
    if eng is None:
        eng = _ENGINES[device] = OverlapEngine(device)
    return eng


_REF_MODULES = {}


def reference_module(name: str):
    """The reference's own module `name` (for the symbols outside the accelerated path), or None.

    Forwarding is opt-in: it happens only when ``OVL_REFERENCE_DIR`` names a checkout of the
    reference.  The reference's functions that call the builder look it up in their module globals
    (overlapGraphs.py:167), so the forwarded module gets the GPU builder / aligner patched in.
    Optional imports of the reference that are not on the accelerated path (Bio, matplotlib) are
    stubbed only while the module executes; sys.modules is left as it was found."""
    import importlib.util
    import os
    import sys
    from unittest.mock import MagicMock
    if name in _REF_MODULES:
        return _REF_MODULES[name]
    ref_dir = os.environ.get("OVL_REFERENCE_DIR")
    if not ref_dir:
        return None
    path = os.path.join(ref_dir, name + ".py")
    mod = None
    if os.path.isfile(path):
        stubbed = []
        for m in ("Bio", "Bio.Align", "matplotlib", "matplotlib.pyplot"):   # not on the hot path, may be absent
            if m in sys.modules:
                continue
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = MagicMock()
                stubbed.append(m)
        spec = importlib.util.spec_from_file_location(f"_ovl_reference_{name}", path)
        mod = importlib.util.module_from_spec(spec)
        saved_path = list(sys.path)
        had = {m: m in sys.modules for m in ("aligners", "overlapGraphs")}
        sys.path.insert(0, ref_dir)
        try:
            spec.loader.exec_module(mod)
        except ImportError:
            mod = None                          # e.g. numba missing: the caller turns this into AttributeError
        finally:
            sys.path[:] = saved_path
            for m, was in had.items():          # do not leave the reference registered under
                if not was:                     # the drop-in's module names
                    sys.modules.pop(m, None)
            for m in stubbed:                   # a later `import matplotlib` must not get a mock
                sys.modules.pop(m, None)
        if mod is not None and name == "overlapGraphs":
            from . import overlapGraphs as dropin
            from . import aligners as dropin_al
            mod.construct_overlap_graph_nx_k = dropin.construct_overlap_graph_nx_k
            mod.remove_cycles_from_graph = dropin.remove_cycles_from_graph      # looked up at call time, :171
            mod.overlap_alignment = dropin_al.overlap_alignment
    _REF_MODULES[name] = mod
    return mod
