"""Generate tests/golden/cycles.json from the LIVE, unmodified reference (build container only):
    python tests/golden/make_golden_cycles.py
For every graph case of graphs.json the reference builds its overlap graph and runs its own
remove_cycles_from_graph (overlapGraphs.py:106-130); the fixture records the sequence of removed edges
(as node indices in insertion order) -- what the drop-in with the device pre-pass must reproduce."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

_, ref_graphs = ref_loader.load()


def main():
    with open(os.path.join(HERE, "graphs.json")) as fh:
        cases = json.load(fh)["cases"]
    out = []
    for c in cases:
        G, _ = ref_graphs.construct_overlap_graph_nx_k(c["reads"], k=c["k"])
        if G.number_of_edges() > 6000:
            continue                                   # dense all-pairs graphs: thousands of find_cycle restarts
        idx = {n: i for i, n in enumerate(G.nodes)}
        removed = []
        orig = G.remove_edge

        def logging_remove(u, v, _orig=orig, _removed=removed, _idx=idx):
            _removed.append([_idx[u], _idx[v]])
            _orig(u, v)

        G.remove_edge = logging_remove
        t0 = time.time()
        ref_graphs.remove_cycles_from_graph(G)
        del G.remove_edge
        out.append({"name": c["name"], "k": c["k"], "removed": removed, "edges_left": G.number_of_edges()})
        print(c["name"], len(removed), "edges removed in", round(time.time() - t0, 2), "s")
    with open(os.path.join(HERE, "cycles.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden_cycles.py",
                   "source": "live reference overlapGraphs.remove_cycles_from_graph on the graphs of graphs.json",
                   "cases": out}, fh, separators=(",", ":"))


if __name__ == "__main__":
    main()
