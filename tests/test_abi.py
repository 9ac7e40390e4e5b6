"""CPU-side checks of the boundary: the library builds, loads and exports what include/ovl.h
declares; host-side argument handling of the drop-ins (no GPU compute here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, load_pkg, has_cuda


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ovl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ovl_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    path = ge.build_library()
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ovl.h but not exported"
    nat = load_pkg("_native")
    assert sorted(nat.EXPORTS) == declared, "ctypes signature table and header disagree"


def test_host_only_entry_points():
    nat = load_pkg("_native")
    assert nat.lib.ovl_version() >= 100
    assert nat.lib.ovl_row_words(1) == 4
    assert nat.lib.ovl_row_words(150) == 12
    assert nat.lib.ovl_row_words(1000) == 64
    assert nat.lib.ovl_index_workspace_bytes(1000) > 0
    out = (ctypes.c_int32 * 3)()
    # default scoring, l=150 -> packed 16-bit kernel, 4 lanes x 38 columns
    assert nat.lib.ovl_overlap_dp_plan(150, 10, -1, -2 ** 31, 0, ctypes.byref(out)) == 0
    assert list(out) == [1, 4, 38]
    assert nat.lib.ovl_overlap_dp_plan(100, 10, -1, -2, 0, ctypes.byref(out)) == 0
    assert list(out) == [1, 4, 25]
    assert nat.lib.ovl_overlap_dp_plan(1000, 10, -1, -2 ** 31, 0, ctypes.byref(out)) == 0
    assert list(out) == [1, 32, 32]
    # scores that do not fit 16 bits fall to the int32 kernel
    assert nat.lib.ovl_overlap_dp_plan(150, 10, -1000000, -2 ** 31, 0, ctypes.byref(out)) == 0
    assert out[0] == 2
    # too long for the wavefront kernels
    assert nat.lib.ovl_overlap_dp_plan(2000, 10, -1, -2 ** 31, 0, ctypes.byref(out)) == 0
    assert list(out) == [1, 32, 76]
    assert nat.lib.ovl_overlap_dp_plan(5000, 10, -1, -2, 0, ctypes.byref(out)) == 0 and out[0] == 3
    assert nat.lib.ovl_overlap_dp_plan(50000, 10, -1, -2, 0, ctypes.byref(out)) == nat.OVL_E_UNSUPPORTED
    assert b"50000" in nat.lib.ovl_last_error()


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    pkg = load_pkg()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.construct_overlap_graph_nx_k(["ACGT", "CGTA"], k=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.overlap_alignment("ACGT", "CGTA")


def test_argument_errors_match_reference():
    pkg = load_pkg()
    with pytest.raises(AssertionError):
        pkg.construct_overlap_graph_nx_k(["ACGT"], k=-1)          # overlapGraphs.py:17
    with pytest.raises(TypeError):
        pkg.overlap_alignment(b"ACGT", "ACGT")                    # reference: Numba TypingError


def test_drop_in_importable_as_top_level_modules():
    """The reference's callers do `from aligners import overlap_alignment` (overlapGraphs.py:2)
    and `from overlapGraphs import ...` (testAssembly.py:3)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import overlapGraphs, aligners; "
            "print(overlapGraphs.construct_overlap_graph_nx_k.__name__, aligners.overlap_alignment.__name__)"
            % os.path.join(ROOT, "genome-assembly-using-overlap-graphs_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    assert out.split() == ["construct_overlap_graph_nx_k", "overlap_alignment"]


def test_synth_generator_is_seeded_and_truncates():
    synth = load_pkg("synth")
    g = synth.phix_like_genome()
    assert len(g) == 5386
    b1, o1 = synth.simulate_reads(g, 500, 100, 0.01, seed=3)
    b2, o2 = synth.simulate_reads(g, 500, 100, 0.01, seed=3)
    assert (b1 == b2).all() and (o1 == o2).all()
    lens = o1[1:] - o1[:-1]
    assert lens.max() == 100 and lens.min() >= 1
    ub, uo, counts, r2u = synth.dedup(b1, o1)
    reads = synth.to_strings(b1, o1)
    d = {}
    for r in reads:
        d[r] = d.get(r, 0) + 1
    assert synth.to_strings(ub, uo) == list(d.keys()) and counts.tolist() == list(d.values())


def test_d2h_chunk_schedule_hides_the_copy():
    """The copy of the edge rows runs behind the DP chunk by chunk (engine.d2h_chunk_cuts): the cuts must tile the
    slice, and for a copy that costs up to 2/3 of the DP only ~1/64 of it may stay exposed."""
    eng = load_pkg("engine")
    assert eng.d2h_chunk_cuts(10) == [0, 64]                                   # short lists: one chunk
    cuts = eng.d2h_chunk_cuts(10 ** 9)
    assert cuts[0] == 0 and cuts[-1] == 64 and all(a < b for a, b in zip(cuts, cuts[1:]))
    capped = eng.d2h_chunk_cuts(64_000_000, chunk_pairs=2_000_000)             # at most 2/64 of the slice per chunk
    assert capped[0] == 0 and capped[-1] == 64 and max(b - a for a, b in zip(capped, capped[1:])) <= 2
    assert set(cuts) <= set(capped) | set(range(65))

    def finish(cuts, c):                       # DP produces 64 units in 64 time units; copying a unit costs c
        t_copy = 0.0
        for a, b in zip(cuts, cuts[1:]):
            t_copy = max(t_copy, float(b)) + c * (b - a)      # a chunk's copy starts when it is computed and the engine is free
        return t_copy

    for c in (0.1, 0.3, 0.5, 0.66):
        assert finish(cuts, c) <= 64 * 1.012 + c, (c, finish(cuts, c))
    assert finish([0, 32, 48, 56, 60, 62, 63, 64], 0.66) > 64 * 1.15               # the first schedule of round 2 at 8 GPUs


def test_numa_cpulist_parser():
    par = load_pkg("parallel")
    assert par._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert par._parse_cpulist("") == set()


def test_graph_from_rows_equals_networkx_edge_by_edge():
    import numpy as np
    """overlapGraphs._graph_from_rows fills the DiGraph's adjacency directly; the result must be indistinguishable from
    add_node / add_edge in the reference's order (nodes, adjacency order, predecessor order, shared attribute dicts)."""
    import networkx as nx
    og = load_pkg("overlapGraphs")
    rng = np.random.default_rng(5)
    uniq = ["ACGT" * 3 + str(i) for i in range(40)]
    counts = rng.integers(1, 4, len(uniq)).astype(np.int32)
    n_nodes = int(counts.sum())
    pairs = sorted({(int(a), int(b)) for a, b in rng.integers(0, n_nodes, (600, 2)) if a != b})
    edges = np.array([(a, b, int(rng.integers(-5, 1500)), int(rng.integers(0, 150))) for a, b in pairs], dtype=np.int32)
    got = og._graph_from_rows(uniq, counts, edges)
    names = [f"{r}_{c}" for r, cnt in zip(uniq, counts.tolist()) for c in range(cnt)]
    want = nx.DiGraph()
    for n in names:
        want.add_node(n)
    for a, b, w, e in edges.tolist():
        want.add_edge(names[a], names[b], weight=w, end_position=e)
    assert list(got.nodes) == list(want.nodes)
    assert list(got.edges(data=True)) == list(want.edges(data=True))
    assert all(list(got.pred[n]) == list(want.pred[n]) for n in names)
    assert got.number_of_edges() == len(pairs) and nx.is_isomorphic(got, want) is True
    u, v = names[pairs[0][0]], names[pairs[0][1]]
    assert got[u][v] is got.pred[v][u]                      # one dict per edge, as NetworkX keeps it
    assert all(type(d["weight"]) is int and type(d["end_position"]) is int for _, _, d in got.edges(data=True))
    got.remove_edge(u, v)                                   # the caller mutates the graph (overlapGraphs.py:117)
    assert not got.has_edge(u, v) and u not in got.pred[v]
    empty = og._graph_from_rows(uniq, counts, np.zeros((0, 4), np.int32))
    assert list(empty.nodes) == names and empty.number_of_edges() == 0
