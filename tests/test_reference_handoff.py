"""Downstream hand-off (build container only: needs the live reference, no GPU).

The GPU builder returns a DiGraph rebuilt from an edge-row array with add_nodes_from /
add_edges_from.  The reference's consumers (cycle removal, topological sort, contig walk --
overlapGraphs.py:106-193) are order sensitive, so this checks that a graph rebuilt that way from
the oracle's edge rows drives the UNMODIFIED reference assembly to exactly the contigs the
reference's own builder gives.  (The GPU edge rows themselves are checked bit-exactly against the
same oracle rows in tests/test_gpu_parity.py.)"""
import contextlib
import io
import random

import numpy as np
import pytest

from conftest import load_pkg
from oracle import overlap_oracle as orc
from oracle import ref_loader

pytestmark = [pytest.mark.live_reference,
              pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present (GPU box)")]


def _rows_from_oracle(reads, k):
    nodes, edges, rc = orc.construct_overlap_graph(reads, k)
    idx = {n: i for i, n in enumerate(nodes)}
    rows = np.array([[idx[u], idx[v], w, e] for u, v, w, e in edges], dtype=np.int32).reshape(-1, 4)
    uniq = list(rc.keys())
    counts = np.fromiter(rc.values(), dtype=np.int32, count=len(uniq))
    return uniq, counts, rows, rc


@pytest.mark.parametrize("k", [5, 10])
def test_rebuilt_graph_gives_identical_contigs(k):
    _, ref_graphs = ref_loader.load()
    og = load_pkg("overlapGraphs")          # importing the drop-in needs neither the GPU nor the .so
    rng = random.Random(4242 + k)
    genome = "".join(rng.choice("ACGT") for _ in range(1500))
    reads = []
    for _ in range(400):
        st = rng.randrange(len(genome))
        r = genome[st:st + 60]
        reads.append("".join(ch if rng.random() > 0.01 else rng.choice("ACGT") for ch in r))
    params = {"experiment_name": "t", "N": len(reads), "l": 60, "error_prob": 0.01, "k": k, "num_iteration": 0}

    with contextlib.redirect_stdout(io.StringIO()):
        want = ref_graphs.assemble_contigs_using_overlap_graphs(list(reads), k=k, params=dict(params))

    def rebuilt_builder(rd, k=5):
        uniq, counts, rows, rc = _rows_from_oracle(rd, k)
        return og._graph_from_rows(uniq, counts, rows), rc

    original = ref_graphs.construct_overlap_graph_nx_k
    ref_graphs.construct_overlap_graph_nx_k = rebuilt_builder       # looked up at call time, overlapGraphs.py:167
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            got = ref_graphs.assemble_contigs_using_overlap_graphs(list(reads), k=k, params=dict(params))
    finally:
        ref_graphs.construct_overlap_graph_nx_k = original
    assert got == want
    assert len(got) > 0


def test_forwarding_module_is_patched_with_the_drop_ins(monkeypatch):
    import sys
    monkeypatch.setenv("OVL_REFERENCE_DIR", ref_loader.REFERENCE_DIR)      # forwarding is opt-in
    eng = load_pkg("engine")
    eng._REF_MODULES.clear()
    had_mpl = "matplotlib" in sys.modules
    og = load_pkg("overlapGraphs")
    al = load_pkg("aligners")
    ref = eng.reference_module("overlapGraphs")
    assert ref is not None
    assert ref.construct_overlap_graph_nx_k is og.construct_overlap_graph_nx_k
    assert ref.overlap_alignment is al.overlap_alignment
    # symbols outside the accelerated path resolve through the drop-in module
    assert ref.remove_cycles_from_graph is og.remove_cycles_from_graph
    assert og.create_contig is ref.create_contig
    assert og.assemble_contigs_using_overlap_graphs is ref.assemble_contigs_using_overlap_graphs
    # dunder probes never reach the forwarding, and the import stubs do not outlive the module load
    assert not hasattr(og, "__wrapped__")
    assert ("matplotlib" in sys.modules) == had_mpl
    eng._REF_MODULES.clear()


def test_no_forwarding_without_opt_in(monkeypatch):
    monkeypatch.delenv("OVL_REFERENCE_DIR", raising=False)
    eng = load_pkg("engine")
    eng._REF_MODULES.clear()
    og = load_pkg("overlapGraphs")
    assert eng.reference_module("overlapGraphs") is None
    with pytest.raises(AttributeError):
        og.create_contig
