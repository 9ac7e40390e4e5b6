"""Drop-in for the reference's ``overlapGraphs.construct_overlap_graph_nx_k``
(overlapGraphs.py:5-61): same signature, same ``(nx.DiGraph, read_copies)`` return value,
same nodes, edges, attributes and iteration order.

What stays on the host is what the return value itself is made of: the ``read_copies`` dict
(overlapGraphs.py:18-20), the node-name strings and the NetworkX container.  The prefix
index, the candidate lookup, the overlap DP and the copy x copy edge expansion
(overlapGraphs.py:30-60) run on the GPU.  Other symbols of the reference module
(assemble_contigs_using_overlap_graphs, ...) are forwarded to the reference checkout named by
``OVL_REFERENCE_DIR`` (opt-in); they call this builder through the module global, as in
overlapGraphs.py:167.

Thread safety: the process-wide engine hands out views of one pinned result buffer, so the
builders hold ``engine.lock`` from the GPU call until the rows are consumed; concurrent calls
from several threads serialise on it.
"""
from __future__ import annotations

import networkx as nx
import numpy as np

try:
    from . import engine as _engine
except ImportError:      # imported as a top-level module with the package directory on sys.path
    import importlib as _il
    import os as _os
    import sys as _sys
    _pkg_dir = _os.path.dirname(_os.path.abspath(__file__))
    if _os.path.dirname(_pkg_dir) not in _sys.path:
        _sys.path.append(_os.path.dirname(_pkg_dir))
    _engine = _il.import_module(_os.path.basename(_pkg_dir) + ".engine")


def _flatten(uniq):
    """Unique reads -> (ASCII bytes, int64 offsets)."""
    lens = np.fromiter((len(r) for r in uniq), dtype=np.int64, count=len(uniq))
    offsets = np.zeros(len(uniq) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    joined = "".join(uniq)
    try:
        raw = joined.encode("latin-1")            # one byte per character, so that offsets are character offsets
    except UnicodeEncodeError as exc:
        raise _engine.nat.OvlUnsupported("reads contain characters above U+00FF; the CUDA builder works on "
                                         "single-byte symbols") from exc
    bases = np.frombuffer(raw, dtype=np.uint8) if raw else np.zeros(0, np.uint8)
    return bases, offsets


def _to_acgt(bases):
    """Re-letter a read set that uses at most four distinct symbols (lower case, RNA, ...) as A/C/G/T.
    Keys and the DP only test symbols for equality, so any bijection leaves every result unchanged.
    Returns None when there are more than four symbols (byte-coded general kernels take over)."""
    symbols = np.unique(bases)
    if symbols.size > 4:
        return None
    lut = np.zeros(256, dtype=np.uint8)
    lut[symbols] = np.frombuffer(b"ACGT", dtype=np.uint8)[:symbols.size]
    return lut[bases]


def overlap_edge_rows(reads, k=5, _reuse_host_buffer=False, min_weight=None):
    """The device part of the builder: returns (read_copies, uniq, counts, edges) where edges
    is int32[E, 4] = (node_a, node_b, weight, end_position) in insertion order and node ids
    number the copies in (uid, copy) order.  min_weight keeps only rows with weight >= it."""
    assert k >= 0, "k-mer length must be non-negative"               # overlapGraphs.py:17
    read_copies, uniq, counts = _dedup(reads)
    if len(uniq) == 0:
        return read_copies, uniq, counts, np.zeros((0, 4), np.int32)
    bases, offsets = _flatten(uniq)
    eng = _engine.get_engine()
    try:
        edges = eng.overlap_edges(bases, offsets, counts, k, reuse_host_buffer=_reuse_host_buffer, min_weight=min_weight)
    except _engine.nat.OvlBadAlphabet:
        # raised right after the pack + count stages, before any pair list is built on the mis-coded rows
        relettered = _to_acgt(bases)
        if relettered is not None:
            edges = eng.overlap_edges(relettered, offsets, counts, k, reuse_host_buffer=_reuse_host_buffer,
                                      min_weight=min_weight)
        else:
            # more than four symbols (e.g. reads with N): byte-coded rows and the general kernels -- the
            # reference compares arbitrary characters (aligners.py:35), so does this path
            edges = eng.overlap_edges(bases, offsets, counts, k, reuse_host_buffer=_reuse_host_buffer,
                                      min_weight=min_weight, code_bits=8)
    return read_copies, uniq, counts, edges


def _graph_from_rows(uniq, counts, edges):
    overlap_graph = nx.DiGraph()
    names = [f"{read}_{c}" for read, cnt in zip(uniq, counts.tolist()) for c in range(cnt)]   # :25-28
    overlap_graph.add_nodes_from(names)
    if edges.shape[0]:
        cols = edges.T.tolist()               # Python ints, like the reference's edge attributes
        succ, pred = getattr(overlap_graph, "_succ", None), getattr(overlap_graph, "_pred", None)
        if (isinstance(succ, dict) and isinstance(pred, dict) and overlap_graph.edge_attr_dict_factory is dict
                and overlap_graph.adjlist_inner_dict_factory is dict):
            # What add_edges_from builds edge by edge -- ONE attribute dict shared by _succ[u][v] and _pred[v][u], rows
            # in insertion order -- without its per-edge membership tests and dict copies (a quarter of the wall time
            # of a 3 M-edge graph).  Every (u, v) occurs once (reads are unique, copies are distinct nodes), so no
            # edge is ever updated, and all nodes exist already.
            srow = [succ[n] for n in names]
            prow = [pred[n] for n in names]
            for u, v, w, e in zip(cols[0], cols[1], cols[2], cols[3]):
                d = {"weight": w, "end_position": e}
                srow[u][names[v]] = d
                prow[v][names[u]] = d
            clear = getattr(nx, "_clear_cache", None)
            if clear is not None:
                clear(overlap_graph)
        else:
            overlap_graph.add_edges_from(
                (names[u], names[v], {"weight": w, "end_position": e})
                for u, v, w, e in zip(cols[0], cols[1], cols[2], cols[3]))
    return overlap_graph


def construct_overlap_graph_nx_k(reads, k=5):
    """Construct the overlap graph -- see overlapGraphs.py:5-16 for the contract."""
    assert k >= 0, "k-mer length must be non-negative"               # overlapGraphs.py:17
    with _engine.get_engine().lock:           # the rows are a view of the engine's pinned buffer: consumed right below
        read_copies, uniq, counts, edges = overlap_edge_rows(reads, k, _reuse_host_buffer=True)
        return _graph_from_rows(uniq, counts, edges), read_copies


def _dedup(reads):
    read_copies = {}
    for read in reads:                                               # overlapGraphs.py:18-20
        read_copies[read] = read_copies.get(read, 0) + 1
    uniq = list(read_copies.keys())
    for r in uniq:
        if not isinstance(r, str):
            raise TypeError("reads must be str")
    counts = np.fromiter(read_copies.values(), dtype=np.int32, count=len(uniq))
    return read_copies, uniq, counts


def overlap_edge_rows_batch(read_lists, k=5):
    """Many independent read sets in ONE GPU job: returns, per read set, (read_copies, uniq, counts, edges)
    exactly as overlap_edge_rows(reads, k) would -- same rows, same order -- but with one pack / index /
    join / DP pass over all sets: every read is tagged with its set number in the key bits above the
    k-mer, so reads of different sets never become candidates.  (Not in the reference: this is the
    graph-build step of its parameter sweep, experiments.py:451-539, which the reference runs as one
    process per parameter set.)  The returned rows are copies (they outlive the next engine call)."""
    assert k >= 0, "k-mer length must be non-negative"
    per_set = [_dedup(reads) for reads in read_lists]
    if k == 0 or len(read_lists) <= 1:
        return [overlap_edge_rows(reads, k) for reads in read_lists]
    all_uniq = [r for _, uniq, _ in per_set for r in uniq]
    if not all_uniq:
        return [(rc, uniq, counts, np.zeros((0, 4), np.int32)) for rc, uniq, counts in per_set]
    counts_all = np.concatenate([c for _, _, c in per_set])
    seg = np.concatenate([np.full(len(u), s_id, dtype=np.int32) for s_id, (_, u, _) in enumerate(per_set)])
    bases, offsets = _flatten(all_uniq)
    has_dups = bool(counts_all.max() > 1)
    eng = _engine.get_engine()
    with eng.lock:
        def job(b):
            # counts are passed when any read repeats, so that node ids are offsets into the concatenated copy list
            return eng.overlap_edges(b, offsets, counts_all if has_dups else None, k, reuse_host_buffer=True,
                                     segments=seg, n_segments=len(read_lists))
        try:
            try:
                edges = job(bases)
            except _engine.nat.OvlBadAlphabet:
                relettered = _to_acgt(bases)             # at most four distinct symbols over ALL sets: exact re-lettering
                if relettered is None:
                    raise
                edges = job(relettered)
        except _engine.nat.OvlUnsupported:
            # what the one-job path does not cover (segment tag + 2k > 64 key bits, more than four symbols):
            # one build per read set
            return [overlap_edge_rows(reads, k) for reads in read_lists]
        node_base = np.zeros(len(per_set) + 1, dtype=np.int64)
        if has_dups:
            np.cumsum([int(c.sum()) for _, _, c in per_set], out=node_base[1:])
        else:
            np.cumsum([len(u) for _, u, _ in per_set], out=node_base[1:])
        bounds = np.searchsorted(edges[:, 0], node_base, side="left") if edges.shape[0] else np.zeros(len(per_set) + 1, np.int64)
        out = []
        for s_id, (read_copies, uniq, counts) in enumerate(per_set):
            rows = edges[bounds[s_id]:bounds[s_id + 1]].copy()
            rows[:, 0] -= int(node_base[s_id])
            rows[:, 1] -= int(node_base[s_id])
            out.append((read_copies, uniq, counts, rows))
    return out


def construct_overlap_graphs_batch(read_lists, k=5):
    """[construct_overlap_graph_nx_k(reads, k) for reads in read_lists] -- same graphs, same order -- from
    ONE GPU job over all read sets (overlap_edge_rows_batch)."""
    return [(_graph_from_rows(uniq, counts, rows), read_copies)
            for read_copies, uniq, counts, rows in overlap_edge_rows_batch(read_lists, k)]


def construct_overlap_graph_string(reads):
    """Drop-in for overlapGraphs.py:196-232: every ordered pair of distinct reads is aligned and
    edges are kept only for score > 0 (same node naming and copy expansion as the k-mer builder)."""
    with _engine.get_engine().lock:
        read_copies, uniq, counts, edges = overlap_edge_rows(reads, 0, _reuse_host_buffer=True, min_weight=1)
        return _graph_from_rows(uniq, counts, edges), read_copies


def construct_string_graph(reads):
    """Drop-in for overlapGraphs.py:332-351: nodes are the reads themselves; for every position pair
    i < j of the READ LIST (duplicates included, itertools.combinations order) the pair
    (reads[i], reads[j]) is aligned and an edge is added when score > 0.

    A pair of distinct strings (u, v) is visited iff some occurrence of u precedes some occurrence
    of v; (u, u) is visited iff u occurs twice.  Each needed pair is aligned once on the GPU; the
    graph is then assembled in the reference's insertion order: sources by first appearance, the
    successors of u by their first occurrence after u's first occurrence."""
    graph = nx.DiGraph()
    first, last, order = {}, {}, []
    for pos, read in enumerate(reads):
        if not isinstance(read, str):
            raise TypeError("reads must be str")
        if read not in first:
            first[read] = pos
            order.append(read)
        last[read] = pos
    graph.add_nodes_from(order)
    U = len(order)
    if U == 0 or len(reads) < 2:
        print(f"graph: {graph.edges}")                               # overlapGraphs.py:350
        return graph
    uid = {r: i for i, r in enumerate(order)}
    pos_uid = np.fromiter((uid[r] for r in reads), dtype=np.int64, count=len(reads))
    first_pos = np.fromiter((first[r] for r in order), dtype=np.int64, count=U)
    last_pos = np.fromiter((last[r] for r in order), dtype=np.int64, count=U)
    pa, pb = [], []
    for u in range(U):
        sub = pos_uid[first_pos[u] + 1:]
        if sub.size == 0:
            continue
        vals, idx = np.unique(sub, return_index=True)
        succ = vals[np.argsort(idx, kind="stable")]                  # first occurrence after u's first occurrence
        pa.append(np.full(succ.shape[0], u, dtype=np.int32))
        pb.append(succ.astype(np.int32))
    pa = np.concatenate(pa) if pa else np.zeros(0, np.int32)
    pb = np.concatenate(pb) if pb else np.zeros(0, np.int32)
    assert np.all((last_pos[pb] > first_pos[pa]))
    if pa.size:
        bases, offsets = _flatten(order)
        eng = _engine.get_engine()
        with eng.lock:
            rows = eng.overlap_edges(bases, offsets, None, 0, pairs=(pa, pb), min_weight=1, reuse_host_buffer=True)
            cols = rows.T.tolist()
        graph.add_edges_from((order[u], order[v], {"weight": w, "end_position": e})
                             for u, v, w, e in zip(cols[0], cols[1], cols[2], cols[3]))
    print(f"graph: {graph.edges}")                                   # overlapGraphs.py:350
    return graph


def remove_cycles_from_graph(overlap_graph):
    """Drop-in for overlapGraphs.py:106-130: remove the weakest edge of the cycle nx.find_cycle reports until
    the graph is a DAG -- the same edges, removed in the same order, as the reference.

    A device pre-pass (ovl_trim_sinks) first peels off every node that cannot reach a cycle.  Such a node is
    never on a reported cycle, and a depth-first search that enters one only backtracks out of it, so the
    search on the surviving nodes (same node order, same adjacency order) reports the same cycles.  The
    reference restarts nx.find_cycle from scratch after every removal; here each restart walks the survivors
    only, and a graph without any cycle costs no search at all."""
    G = overlap_graph
    nodes = list(G.nodes)
    n = len(nodes)
    if n == 0 or G.number_of_edges() == 0:
        return G
    idx = {v: i for i, v in enumerate(nodes)}
    edges = list(G.edges)                                    # (source order, adjacency order)
    src = np.fromiter((idx[u] for u, _ in edges), dtype=np.int32, count=len(edges))
    dst = np.fromiter((idx[v] for _, v in edges), dtype=np.int32, count=len(edges))
    loops = src == dst
    keep, _ = _engine.get_engine().trim_sinks(src[~loops], dst[~loops], n)
    keep[src[loops]] = True                                  # a self loop is a cycle of its own
    if loops.any():
        # nodes that reach a self loop must stay too: fall back to the plain search on the whole graph
        keep[:] = True
    if not keep.any():
        return G                                             # already a DAG
    H = nx.DiGraph()
    H.add_nodes_from(v for v, k in zip(nodes, keep.tolist()) if k)
    H.add_edges_from((u, v) for (u, v), a, b in zip(edges, keep[src].tolist(), keep[dst].tolist()) if a and b)
    while True:
        try:
            cycle = nx.find_cycle(H, orientation='original')
        except nx.NetworkXNoCycle:
            break
        u, v, _w = min(((u, v, G[u][v]["weight"]) for u, v, _ in cycle), key=lambda x: x[2])   # overlapGraphs.py:126
        G.remove_edge(u, v)
        H.remove_edge(u, v)
    return G


def __getattr__(name):
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    ref = _engine.reference_module("overlapGraphs")
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    raise AttributeError(f"module 'overlapGraphs' (B200 drop-in) has no attribute {name!r}")
