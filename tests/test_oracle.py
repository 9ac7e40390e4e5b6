"""The oracle against the golden vectors produced by the live reference (CPU only)."""
import hashlib
import random

import numpy as np
import pytest

from oracle import overlap_oracle as orc
from oracle import ref_loader


def test_pairs_match_reference_golden(golden_pairs):
    assert len(golden_pairs) >= 400
    for c in golden_pairs:
        got = orc.overlap_alignment(c["s"], c["t"], c["match"], c["mismatch"], c["indel"])
        want = (c["to_print"], c["align_s"], c["align_t"], c["score"], c["end"])
        assert got == want, (c["s"], c["t"], c["match"], c["mismatch"], c["indel"])
        assert type(got[3]) is int and type(got[4]) is int


def test_rolling_variant_equals_full(golden_pairs):
    reads = []
    for c in golden_pairs:
        reads += [c["s"], c["t"]]
    bases, off = orc.concat_reads(reads)
    pa = np.arange(0, len(reads), 2, dtype=np.int32)
    pb = pa + 1
    for prm in [(10, -1, -2 ** 31), (10, -1, -2), (2, -3, -2), (10, -1, 0)]:
        s1, e1 = orc.overlap_pairs(bases, off, pa, pb, *prm, full=True)
        s2, e2 = orc.overlap_pairs(bases, off, pa, pb, *prm, full=False, nthreads=2)
        assert np.array_equal(s1, s2) and np.array_equal(e1, e2)


def test_gapless_closed_form():
    """with the default indel the last row is a sum over a diagonal."""
    rng = random.Random(5)
    for _ in range(200):
        s = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 40)))
        t = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 40)))
        n, m = len(s), len(t)
        best, bj = 0, 0
        for j in range(1, m + 1):
            v = sum(10 if s[n - 1 - d] == t[j - 1 - d] else -1 for d in range(min(n, j)))
            if v > best:
                best, bj = v, j
        _, a_s, a_t, score, end = orc.overlap_alignment(s, t)
        assert (score, end) == (best, bj)
        L = min(n, end)
        assert a_s == s[n - L:] and a_t == t[end - L:end]


def test_graphs_match_reference_golden(golden_graphs):
    assert len(golden_graphs) >= 20
    for c in golden_graphs:
        nodes, edges, read_copies = orc.construct_overlap_graph(c["reads"], c["k"])
        assert [[r, n] for r, n in read_copies.items()] == c["read_copies"], c["name"]
        assert len(nodes) == c["n_nodes"]
        assert hashlib.sha256("\n".join(nodes).encode()).hexdigest() == c["nodes_sha256"], c["name"]
        idx = {n: i for i, n in enumerate(nodes)}
        got = [[idx[u], idx[v], w, e] for u, v, w, e in edges]
        # the fixture lists G.edges(data=True): NetworkX iterates by source node, then by
        # insertion order, i.e. a stable sort of the insertion order by source node
        got.sort(key=lambda r: r[0])
        assert got == c["edges"], c["name"]


def test_allpairs_builders_match_reference_golden(golden_allpairs):
    """overlapGraphs.construct_overlap_graph_string (:196-232) and construct_string_graph (:332-351)."""
    for c in golden_allpairs:
        nodes, edges, rc = orc.construct_overlap_graph_string(c["reads"])
        assert nodes == c["string_nodes"], c["name"]
        assert [[r, n] for r, n in rc.items()] == c["string_read_copies"]
        idx = {n: i for i, n in enumerate(nodes)}
        got = sorted(([idx[u], idx[v], w, e] for u, v, w, e in edges), key=lambda r: r[0])
        assert got == c["string_edges"], c["name"]
        H = orc.construct_string_graph(c["reads"])
        hn = list(H.nodes)
        assert hn == c["sg_nodes"], c["name"]
        hidx = {n: i for i, n in enumerate(hn)}
        assert [[hidx[u], hidx[v], d["weight"], d["end_position"]] for u, v, d in H.edges(data=True)] == c["sg_edges"]
        assert [[hidx[p] for p in H.pred[n]] for n in hn] == c["sg_pred"]


def test_local_alignment_matches_reference_golden(golden_local):
    """aligners.local_alignment (aligners.py:85-167)."""
    assert len(golden_local["local"]) >= 150
    for c in golden_local["local"]:
        got = orc.local_alignment(c["query"], c["reference"], c["match"], c["mismatch"], c["indel"])
        assert list(got) == c["out"], (c["query"], c["reference"], c["match"], c["mismatch"], c["indel"])


def test_negative_k_asserts():
    with pytest.raises(AssertionError):
        orc.construct_overlap_graph(["ACGT"], k=-1)


@pytest.mark.live_reference
@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present (GPU box)")
def test_oracle_vs_live_reference_random():
    ref_aligners, ref_graphs = ref_loader.load()
    rng = random.Random(77)
    for it in range(150):
        s = "".join(rng.choice("ACGT") for _ in range(rng.randint(0, 50)))
        t = "".join(rng.choice("ACGT") for _ in range(rng.randint(0, 50)))
        if it % 2:
            assert orc.overlap_alignment(s, t) == ref_aligners.overlap_alignment(s, t)
        else:
            assert orc.overlap_alignment(s, t, 10, -1, -2) == ref_aligners.overlap_alignment(s, t, 10, -1, -2)
    g = "".join(rng.choice("ACGT") for _ in range(300))
    reads = []
    for _ in range(400):
        st = rng.randrange(len(g))
        reads.append(g[st:st + 25])
    for k in (4, 6):
        G, rc = ref_graphs.construct_overlap_graph_nx_k(reads, k=k)
        nodes, edges, rc2 = orc.construct_overlap_graph(reads, k)
        assert list(rc.items()) == list(rc2.items())
        assert list(G.nodes) == nodes
        G2 = orc.to_networkx(nodes, edges)
        assert list(G.edges(data=True)) == list(G2.edges(data=True))
        assert [list(G.pred[n]) for n in G.nodes] == [list(G2.pred[n]) for n in G2.nodes]


def test_oracle_cycle_removal_matches_live_reference_fixture():
    """oracle.remove_cycles_from_graph (the literal restatement of overlapGraphs.py:106-130) on the golden graphs
    reproduces the removed-edge sequences recorded from the live reference (tests/golden/make_golden_cycles.py)."""
    import json
    import os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "graphs.json")) as fh:
        graphs = {c["name"]: c for c in json.load(fh)["cases"]}
    with open(os.path.join(GOLDEN, "cycles.json")) as fh:
        cycles = json.load(fh)["cases"]
    total = 0
    for cyc in cycles:
        c = graphs[cyc["name"]]
        nodes, edges, _ = orc.construct_overlap_graph(c["reads"], c["k"])
        G = orc.to_networkx(nodes, edges)
        idx = {v: i for i, v in enumerate(nodes)}
        removed = [[idx[u], idx[v]] for u, v in orc.remove_cycles_from_graph(G)]
        assert removed == cyc["removed"], cyc["name"]
        assert G.number_of_edges() == cyc["edges_left"]
        total += len(removed)
    assert total > 1500
