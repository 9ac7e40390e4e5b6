"""Python face of the parity oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package never does.

It wraps ``overlap_oracle.c`` (the C restatement of ``aligners.py:27-82``) and restates
the graph builder ``overlapGraphs.py:5-61`` on top of it.  Parity pinning: the reference
holds no golden vectors for this path, so the oracle is pinned against outputs of the
live, unmodified reference generated in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.json``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboverlap_oracle.so")
_lib = None

INDEL_DEFAULT = -2 ** 31  # aligners.py:7


def build(force: bool = False) -> str:
    """Compile the C oracle with the Makefile beside it (gcc, OpenMP)."""
    src = os.path.join(_HERE, "overlap_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboverlap_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.ovo_overlap_alignment.argtypes = [u8p, ctypes.c_int32, u8p, ctypes.c_int32,
                                              ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                              i32p, i32p, u8p, u8p, i32p]
        lib.ovo_overlap_alignment.restype = ctypes.c_int
        lib.ovo_overlap_score.argtypes = [u8p, ctypes.c_int32, u8p, ctypes.c_int32,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          i32p, i32p]
        lib.ovo_overlap_score.restype = ctypes.c_int
        lib.ovo_overlap_pairs.argtypes = [u8p, i64p, i32p, i32p, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          i32p, i32p, ctypes.c_int32, ctypes.c_int32]
        lib.ovo_overlap_pairs.restype = ctypes.c_int
        lib.ovo_local_alignment.argtypes = [u8p, ctypes.c_int32, u8p, ctypes.c_int32,
                                            ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                            i32p, u8p, u8p, i32p]
        lib.ovo_local_alignment.restype = ctypes.c_int
        lib.ovo_max_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(_load().ovo_max_threads())


def _as_u8(x) -> np.ndarray:
    if isinstance(x, str):
        # the reference compares unicode code points; the oracle is byte based and is
        # only ever fed latin-1 text (reads are ACGT)
        x = x.encode("latin-1")
    return np.frombuffer(bytes(x), dtype=np.uint8)


def _ptr(a: np.ndarray, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def overlap_alignment(s: str, t: str, match_score: int = 10, mismatch: int = -1,
                      indel: int = INDEL_DEFAULT) -> Tuple[str, str, str, int, int]:
    """Restates ``aligners.overlap_alignment`` (aligners.py:6-82): same 5-tuple."""
    lib = _load()
    sb, tb = _as_u8(s), _as_u8(t)
    n, m = len(sb), len(tb)
    a_s = np.zeros(n + m + 2, dtype=np.uint8)
    a_t = np.zeros(n + m + 2, dtype=np.uint8)
    score, end, L = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    sb_ = sb if n else np.zeros(1, np.uint8)
    tb_ = tb if m else np.zeros(1, np.uint8)
    rc = lib.ovo_overlap_alignment(_ptr(sb_, ctypes.c_uint8), n, _ptr(tb_, ctypes.c_uint8), m,
                                   int(match_score), int(mismatch), int(indel),
                                   ctypes.byref(score), ctypes.byref(end),
                                   _ptr(a_s, ctypes.c_uint8), _ptr(a_t, ctypes.c_uint8),
                                   ctypes.byref(L))
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    align_s = a_s[:L.value].tobytes().decode("latin-1")
    align_t = a_t[:L.value].tobytes().decode("latin-1")
    # aligners.py:78
    to_print = f"\nTarget:   {align_t}\n          {'|' * len(align_t)}\nQuery:    {align_s}"
    return to_print, align_s, align_t, int(score.value), int(end.value)


def local_alignment(query: str, reference: str, match_score: int = 10, mismatch: int = -1, indel: int = -1):
    """Restates ``aligners.local_alignment`` (aligners.py:85-167): same 6-tuple."""
    lib = _load()
    qb, rb = _as_u8(query), _as_u8(reference)
    n, m = len(qb), len(rb)
    a_q = np.zeros(n + m + 2, dtype=np.uint8)
    a_r = np.zeros(n + m + 2, dtype=np.uint8)
    out = np.zeros(4, dtype=np.int32)
    L = ctypes.c_int32()
    qb_ = qb if n else np.zeros(1, np.uint8)
    rb_ = rb if m else np.zeros(1, np.uint8)
    rc = lib.ovo_local_alignment(_ptr(qb_, ctypes.c_uint8), n, _ptr(rb_, ctypes.c_uint8), m,
                                 int(match_score), int(mismatch), int(indel), _ptr(out, ctypes.c_int32),
                                 _ptr(a_q, ctypes.c_uint8), _ptr(a_r, ctypes.c_uint8), ctypes.byref(L))
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    aligned_query = a_q[:L.value].tobytes().decode("latin-1")
    aligned_reference = a_r[:L.value].tobytes().decode("latin-1")
    to_print = (f"\nTarget:   {aligned_reference}\n          {'|' * len(aligned_reference)}\nQuery:    "
                f"{aligned_query}")                                    # aligners.py:166-167
    return to_print, aligned_reference, aligned_query, int(out[0]), int(out[1]), int(out[2])


def concat_reads(reads: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenate reads into one byte array + int64 offsets[len(reads)+1]."""
    lens = np.fromiter((len(r) for r in reads), dtype=np.int64, count=len(reads))
    offsets = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    bases = np.frombuffer("".join(reads).encode("latin-1"), dtype=np.uint8)
    if bases.size == 0:
        bases = np.zeros(1, dtype=np.uint8)
    return bases, offsets


def overlap_pairs(bases: np.ndarray, offsets: np.ndarray, pair_a: np.ndarray, pair_b: np.ndarray,
                  match_score: int = 10, mismatch: int = -1, indel: int = INDEL_DEFAULT,
                  full: bool = False, nthreads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """(score[p], end[p]) of overlap_alignment(reads[a[p]], reads[b[p]]) -- the call site
    overlapGraphs.py:53 applied to an index-pair list.  ``full=True`` does the whole
    reference computation per call (matrices + traceback walk); it is the timed CPU arm."""
    lib = _load()
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    pair_a = np.ascontiguousarray(pair_a, dtype=np.int32)
    pair_b = np.ascontiguousarray(pair_b, dtype=np.int32)
    P = int(pair_a.shape[0])
    score = np.zeros(max(P, 1), dtype=np.int32)
    end = np.zeros(max(P, 1), dtype=np.int32)
    if P:
        rc = lib.ovo_overlap_pairs(_ptr(bases, ctypes.c_uint8), _ptr(offsets, ctypes.c_int64),
                                   _ptr(pair_a, ctypes.c_int32), _ptr(pair_b, ctypes.c_int32), P,
                                   int(match_score), int(mismatch), int(indel),
                                   _ptr(score, ctypes.c_int32), _ptr(end, ctypes.c_int32),
                                   1 if full else 0, int(nthreads))
        if rc != 0:
            raise MemoryError("oracle allocation failed")
    return score[:P], end[:P]


def dedup_reads(reads: Sequence[str]) -> Dict[str, int]:
    """overlapGraphs.py:18-20 -- unique reads in first-appearance order -> multiplicity."""
    read_copies: Dict[str, int] = {}
    for r in reads:
        read_copies[r] = read_copies.get(r, 0) + 1
    return read_copies


def candidate_pairs(uniq: Sequence[str], k: int) -> Tuple[np.ndarray, np.ndarray]:
    """overlapGraphs.py:30-52 on unique-read indices: ordered (a, b) candidate list in the
    reference's visiting order (a ascending, b in bucket-append = ascending order)."""
    assert k >= 0, "k-mer length must be non-negative"          # overlapGraphs.py:17
    U = len(uniq)
    pa: List[int] = []
    pb: List[int] = []
    if k == 0:                                                   # overlapGraphs.py:49
        for a in range(U):
            for b in range(U):
                if a != b:
                    pa.append(a)
                    pb.append(b)
    else:
        index: Dict[str, List[int]] = {}
        for b, r in enumerate(uniq):                             # overlapGraphs.py:33-40
            key = r[:k] if len(r) >= k else r
            index.setdefault(key, []).append(b)
        for a, r in enumerate(uniq):                             # overlapGraphs.py:43-52
            key = r[-k:] if len(r) >= k else r
            for b in index.get(key, ()):
                if b != a:        # unique strings: read_a != read_b  <=>  a != b
                    pa.append(a)
                    pb.append(b)
    return np.asarray(pa, dtype=np.int32), np.asarray(pb, dtype=np.int32)


def construct_overlap_graph(reads: Sequence[str], k: int = 5, nthreads: int = 0, full: bool = False):
    """Restates construct_overlap_graph_nx_k (overlapGraphs.py:5-61) without NetworkX.

    Returns (nodes, edges, read_copies): ``nodes`` in insertion order (uid, copy),
    ``edges`` as (u_name, v_name, weight, end_position) in insertion order
    (a_uid, b_uid, copy_a, copy_b), ``read_copies`` the ordered multiplicity dict.
    """
    assert k >= 0, "k-mer length must be non-negative"
    read_copies = dedup_reads(reads)
    uniq = list(read_copies.keys())
    counts = list(read_copies.values())
    nodes = [f"{r}_{c}" for r, cnt in zip(uniq, counts) for c in range(cnt)]   # :25-28
    pa, pb = candidate_pairs(uniq, k)
    bases, offsets = concat_reads(uniq)
    score, end = overlap_pairs(bases, offsets, pa, pb, full=full, nthreads=nthreads)  # :53 defaults
    edges = []
    for a, b, w, e in zip(pa.tolist(), pb.tolist(), score.tolist(), end.tolist()):
        ra, rb = uniq[a], uniq[b]
        for ca in range(counts[a]):                              # overlapGraphs.py:55-60
            for cb in range(counts[b]):
                edges.append((f"{ra}_{ca}", f"{rb}_{cb}", w, e))
    return nodes, edges, read_copies


def construct_overlap_graph_string(reads: Sequence[str], nthreads: int = 0):
    """Restates overlapGraphs.py:196-232: all ordered pairs of distinct reads, edges for score > 0."""
    nodes, edges, read_copies = construct_overlap_graph(reads, 0, nthreads=nthreads)
    return nodes, [e for e in edges if e[2] > 0], read_copies


def construct_string_graph(reads: Sequence[str]):
    """Restates overlapGraphs.py:332-351 literally: combinations over the read list, score > 0.
    Returns a NetworkX DiGraph built with the same call sequence as the reference."""
    import itertools
    import networkx as nx
    g = nx.DiGraph()
    for r in reads:
        g.add_node(r)
    cache: Dict[Tuple[str, str], Tuple[int, int]] = {}
    for a, b in itertools.combinations(reads, 2):
        if (a, b) not in cache:
            out = overlap_alignment(a, b)
            cache[(a, b)] = (out[3], out[4])
        score, end = cache[(a, b)]
        if score > 0:
            g.add_edge(a, b, weight=score, end_position=end)
    return g


def remove_cycles_from_graph(G):
    """Restates overlapGraphs.py:106-130 literally (NetworkX calls and all); returns the removed edges in order."""
    import networkx as nx
    removed = []
    while True:
        try:
            cycle = nx.find_cycle(G, orientation='original')
        except nx.NetworkXNoCycle:
            break
        u, v, _w = min(((u, v, G[u][v]["weight"]) for u, v, _ in cycle), key=lambda x: x[2])
        G.remove_edge(u, v)
        removed.append((u, v))
    return removed


def to_networkx(nodes, edges):
    import networkx as nx
    g = nx.DiGraph()
    g.add_nodes_from(nodes)
    g.add_edges_from((u, v, {"weight": w, "end_position": e}) for u, v, w, e in edges)
    return g
