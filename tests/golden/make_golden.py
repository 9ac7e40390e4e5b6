"""Generate tests/golden/*.json from the LIVE, unmodified reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference holds no golden vectors of its own for overlap_alignment /
construct_overlap_graph_nx_k, so these fixtures -- outputs of the
reference's own Numba / NetworkX code on seeded inputs -- are what pins the oracle and
the CUDA path.  Inputs are stored alongside outputs; nothing here is needed at test time
except the JSON files.
"""
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

ref_aligners, ref_graphs = ref_loader.load()


def pair_case(s, t, params):
    if params is None:
        out = ref_aligners.overlap_alignment(s, t)
        ma, mi, ind = 10, -1, -2 ** 31
    else:
        ma, mi, ind = params
        out = ref_aligners.overlap_alignment(s, t, ma, mi, ind)
    to_print, a_s, a_t, score, end = out
    assert type(score) is int and type(end) is int
    return {"s": s, "t": t, "match": ma, "mismatch": mi, "indel": ind,
            "to_print": to_print, "align_s": a_s, "align_t": a_t, "score": score, "end": end}


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, p, indel_p=0.0):
    out = []
    for ch in s:
        r = rng.random()
        if r < indel_p / 2:
            continue
        if r < indel_p:
            out.append(rng.choice("ACGT"))
        if rng.random() < p:
            ch = rng.choice([c for c in "ACGT" if c != ch])
        out.append(ch)
    return "".join(out)


def make_pairs():
    rng = random.Random(20261018)
    cases = []
    # known-answer inputs recorded in SURVEY.md section 4
    kat = [("ACGTACGTAA", "GTAACCCC"), ("AAAA", "CCCC"), ("ACG", "ACGTTT"), ("TTACG", "ACG"),
           ("ACGT", "TTACGTGG"), ("A", ""), ("", "T"), ("", ""), ("A", "A"), ("A", "C"),
           ("ACGT", "ACGT"), ("AAAAAAAA", "AAAA"), ("AAAA", "AAAAAAAA"), ("ACACACAC", "CACACACA")]
    for s, t in kat:
        cases.append(pair_case(s, t, None))
    finite = [(10, -1, -2), (10, -1, -11), (10, -1, -1), (1, -1, -1), (5, -4, -3), (2, -3, -2),
              (10, -1, -1000), (10, -1000000, -2 ** 31), (3, 0, -1), (10, -1, 0)]
    for s, t in [("ACGTTACG", "ACGACGGG"), ("ACGTTACG", "TACGGG"), ("AAAACCCCGGGG", "CCCGGGGTT")]:
        for prm in finite:
            cases.append(pair_case(s, t, prm))
    # random / adversarial pairs
    for it in range(420):
        kind = it % 7
        if kind == 0:      # unrelated
            s, t = rand_seq(rng, rng.randint(1, 60)), rand_seq(rng, rng.randint(1, 60))
        elif kind == 1:    # planted exact suffix/prefix overlap
            ov = rand_seq(rng, rng.randint(1, 30))
            s, t = rand_seq(rng, rng.randint(0, 30)) + ov, ov + rand_seq(rng, rng.randint(0, 30))
        elif kind == 2:    # planted overlap with substitutions
            ov = rand_seq(rng, rng.randint(5, 40))
            s = rand_seq(rng, rng.randint(0, 20)) + ov
            t = mutate(rng, ov, 0.1) + rand_seq(rng, rng.randint(0, 20))
        elif kind == 3:    # planted overlap with indels (matters for finite indel)
            ov = rand_seq(rng, rng.randint(8, 40))
            s = rand_seq(rng, rng.randint(0, 20)) + ov
            t = mutate(rng, ov, 0.03, 0.1) + rand_seq(rng, rng.randint(0, 20))
        elif kind == 4:    # low-complexity: many ties
            s, t = rand_seq(rng, rng.randint(1, 50), "AC"), rand_seq(rng, rng.randint(1, 50), "AC")
        elif kind == 5:    # homopolymers / n != m extremes
            s, t = "A" * rng.randint(1, 40), "A" * rng.randint(1, 40)
            if rng.random() < 0.5:
                t = t[:-1] + "C"
        else:              # t contains s (end > len(s) possible)
            s = rand_seq(rng, rng.randint(1, 20))
            t = rand_seq(rng, rng.randint(0, 10)) + s + rand_seq(rng, rng.randint(0, 10))
        prm = None if it % 3 == 0 else finite[rng.randrange(len(finite))]
        cases.append(pair_case(s, t, prm))
    # a few read-sized pairs (l = 100 / 150 / 300)
    for l in (100, 150, 300):
        g = rand_seq(rng, 3 * l)
        for _ in range(6):
            a0 = rng.randint(0, l)
            b0 = a0 + rng.randint(1, l - 1)
            s, t = mutate(rng, g[a0:a0 + l], 0.02), mutate(rng, g[b0:b0 + l], 0.02)
            cases.append(pair_case(s, t, None))
            cases.append(pair_case(s, t, (10, -1, -2)))
    return cases


def graph_case(name, reads, k):
    G, read_copies = ref_graphs.construct_overlap_graph_nx_k(list(reads), k=k)
    nodes = list(G.nodes)
    idx = {n: i for i, n in enumerate(nodes)}
    edges = []
    for u, v, d in G.edges(data=True):
        assert type(d["weight"]) is int and type(d["end_position"]) is int
        edges.append([idx[u], idx[v], d["weight"], d["end_position"]])
    derived = [f"{r}_{c}" for r, cnt in read_copies.items() for c in range(cnt)]
    assert derived == nodes, "node order is (uid, copy) -- overlapGraphs.py:25-28"
    return {"name": name, "k": k, "reads": list(reads),
            "read_copies": [[r, c] for r, c in read_copies.items()],
            "n_nodes": len(nodes),
            "nodes_sha256": hashlib.sha256("\n".join(nodes).encode()).hexdigest(),
            "edges": edges}


def sim_reads(rng, genome, n, l, p):
    # generateErrorFreeReads.py:38-50 + generateErrorProneReads.py:17-28, seeded
    reads = []
    for _ in range(n):
        st = rng.randrange(len(genome))
        r = genome[st:st + l]
        reads.append(mutate(rng, r, p))
    return reads


def make_graphs():
    rng = random.Random(987654321)
    cases = []
    toy = ['TGTTC', 'TGCGT', 'ACGTG', 'CACGT', 'AGCAC', 'GATAG', 'CGATA', 'GTACG', 'CGTAC', 'ATGCG']  # overlapGraphs.py:425
    for k in (0, 1, 2, 3, 5, 6):
        cases.append(graph_case(f"toy_k{k}", toy, k))
    # duplicates + truncated reads + reads shorter than k, tiny genome
    g = rand_seq(rng, 120)
    reads = sim_reads(rng, g, 150, 12, 0.02) + ["ACG", "AC", "A", "ACG", "ACGTACGTACGT", "ACGTACGTACGT"]
    for k in (0, 3, 4, 5, 8, 12, 13):
        cases.append(graph_case(f"dups_k{k}", reads, k))
    # low-complexity genome: big buckets, many self-matching keys (prefix == suffix k-mer)
    g = "ACAC" * 20 + rand_seq(rng, 40) + "AAAAAAAAAAAAAAAA"
    reads = sim_reads(rng, g, 120, 16, 0.0)
    for k in (2, 4, 7):
        cases.append(graph_case(f"lowcomplexity_k{k}", reads, k))
    # k up to 32 and reads of exactly k bases
    g = rand_seq(rng, 400)
    reads = sim_reads(rng, g, 200, 40, 0.0)
    for k in (16, 31, 32, 40):
        cases.append(graph_case(f"longk_k{k}", reads, k))
    # BASELINE.json configs[0]: PhiX174, N=1000, l=100, p=0.01
    fasta = os.path.join(ref_loader.REFERENCE_DIR, "sequence.fasta")
    genome = "".join(line.strip() for line in open(fasta) if not line.startswith(">"))
    reads = sim_reads(rng, genome, 1000, 100, 0.01)
    for k in (5, 10, 15):
        cases.append(graph_case(f"phix_n1000_l100_k{k}", reads, k))
    return cases


def make_local():
    """aligners.local_alignment (aligners.py:85-167) and align_read_or_contig_to_reference (:170-202)."""
    rng = random.Random(8086)
    cases = []

    def one(q, r, prm):
        out = ref_aligners.local_alignment(q, r) if prm is None else ref_aligners.local_alignment(q, r, *prm)
        ma, mi, ind = (10, -1, -1) if prm is None else prm
        assert type(out[3]) is int and type(out[4]) is int and type(out[5]) is int
        cases.append({"query": q, "reference": r, "match": ma, "mismatch": mi, "indel": ind, "out": list(out)})

    for q, r in [("ACGT", "TTACGTGG"), ("AAAA", "CCCC"), ("", "ACGT"), ("ACGT", ""), ("", ""), ("A", "A"),
                 ("ACGTACGT", "ACGTACGT"), ("ACGTTTACGT", "ACGTACGT"), ("GGGACGTGGG", "TTTACGTTTT")]:
        one(q, r, None)
    params = [None, (10, -1, -1), (10, -1, -2), (1, -1, -1), (5, -4, -3), (2, -3, -2), (10, -1, 0), (3, 0, -1)]
    for it in range(160):
        g = rand_seq(rng, rng.randint(20, 120))
        st = rng.randrange(len(g))
        q = mutate(rng, g[st:st + rng.randint(1, 40)], 0.08, 0.08) or "A"
        kind = it % 4
        if kind == 1:
            q = rand_seq(rng, rng.randint(1, 30))
        elif kind == 2:
            q, g = rand_seq(rng, rng.randint(1, 25), "AC"), rand_seq(rng, rng.randint(1, 60), "AC")
        elif kind == 3:
            q = g[st:st + 30] + rand_seq(rng, 5) + g[st:st + 10]
        one(q, g, params[it % len(params)])
    wrapped = []
    genome = rand_seq(rng, 400)
    for _ in range(30):
        st = rng.randrange(len(genome))
        L = rng.choice([5, 12, 30, 60])
        read = mutate(rng, genome[st:st + L], 0.05)
        rl = rng.choice([10, 30, 50])
        out = ref_aligners.align_read_or_contig_to_reference(read, genome, rl)
        wrapped.append({"seq": read, "genome": genome, "read_length": rl, "out": list(out)})
    return {"local": cases, "wrapped": wrapped}


def make_allpairs():
    """overlapGraphs.construct_overlap_graph_string (:196-232) and construct_string_graph (:332-351)."""
    import contextlib
    import io
    rng = random.Random(5150)
    out = []
    toy = ['TGTTC', 'TGCGT', 'ACGTG', 'CACGT', 'AGCAC', 'GATAG', 'CGATA', 'GTACG', 'CGTAC', 'ATGCG']
    g = rand_seq(rng, 90)
    dup_reads = sim_reads(rng, g, 60, 14, 0.03) + ["ACG", "ACG", "A", "TTTTTTTT", "TTTTTTTT", "GGGG", "CCCC"]
    for name, reads in [("toy", toy), ("dups", dup_reads), ("two", ["ACGT", "ACGT"]), ("one", ["ACGT"]), ("empty", [])]:
        G, rc = ref_graphs.construct_overlap_graph_string(list(reads))
        nodes = list(G.nodes)
        idx = {n: i for i, n in enumerate(nodes)}
        case = {"name": name, "reads": list(reads), "string_nodes": nodes,
                "string_read_copies": [[r, c] for r, c in rc.items()],
                "string_edges": [[idx[u], idx[v], d["weight"], d["end_position"]] for u, v, d in G.edges(data=True)]}
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            H = ref_graphs.construct_string_graph(list(reads))
        hn = list(H.nodes)
        hidx = {n: i for i, n in enumerate(hn)}
        case["sg_nodes"] = hn
        case["sg_edges"] = [[hidx[u], hidx[v], d["weight"], d["end_position"]] for u, v, d in H.edges(data=True)]
        case["sg_pred"] = [[hidx[p] for p in H.pred[n]] for n in hn]
        case["sg_stdout_sha256"] = hashlib.sha256(buf.getvalue().encode()).hexdigest()
        out.append(case)
    return out


if __name__ == "__main__":
    pairs = make_pairs()
    with open(os.path.join(HERE, "pairs.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py", "source": "live reference aligners.overlap_alignment",
                   "cases": pairs}, fh, separators=(",", ":"))
    graphs = make_graphs()
    with open(os.path.join(HERE, "graphs.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py",
                   "source": "live reference overlapGraphs.construct_overlap_graph_nx_k",
                   "cases": graphs}, fh, separators=(",", ":"))
    loc = make_local()
    with open(os.path.join(HERE, "local.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py",
                   "source": "live reference aligners.local_alignment / align_read_or_contig_to_reference", **loc},
                  fh, separators=(",", ":"))
    print(len(loc["local"]), "local-alignment cases;", len(loc["wrapped"]), "wrapper cases")
    allp = make_allpairs()
    with open(os.path.join(HERE, "allpairs.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py",
                   "source": "live reference overlapGraphs.construct_overlap_graph_string / construct_string_graph",
                   "cases": allp}, fh, separators=(",", ":"))
    print(len(allp), "all-pairs cases")
    print(len(pairs), "pair cases;", len(graphs), "graph cases;",
          sum(len(c["edges"]) for c in graphs), "edges")
