"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden
fixtures produced by the live reference.  Bit-exact: integer scores, end positions, edges."""
import hashlib
import random

import numpy as np
import pytest

from conftest import load_pkg, has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]

from oracle import overlap_oracle as orc  # noqa: E402  (the checker)


@pytest.fixture(scope="module")
def eng():
    return load_pkg("engine").get_engine()


@pytest.fixture(scope="module")
def nat():
    return load_pkg("_native")


def rand_reads(rng, n, lo, hi, alphabet="ACGT"):
    return ["".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi))) for _ in range(n)]


def overlapping_reads(rng, n, genome_len, read_len, p):
    g = "".join(rng.choice("ACGT") for _ in range(genome_len))
    out = []
    for _ in range(n):
        st = rng.randrange(genome_len)
        r = g[st:st + read_len]
        r = "".join(ch if rng.random() > p else rng.choice([c for c in "ACGT" if c != ch]) for ch in r)
        out.append(r)
    return out


def upload(eng, reads):
    bases, offsets = orc.concat_reads(reads)
    return eng.upload_reads(bases[:int(offsets[-1])], offsets), bases, offsets


# --------------------------------------------------------------------------- K7 single pair
def test_single_pair_golden(golden_pairs):
    al = load_pkg("aligners")
    for c in golden_pairs:
        got = al.overlap_alignment(c["s"], c["t"], c["match"], c["mismatch"], c["indel"])
        want = (c["to_print"], c["align_s"], c["align_t"], c["score"], c["end"])
        assert got == want, (c["s"], c["t"], c["match"], c["mismatch"], c["indel"])
        assert type(got[3]) is int and type(got[4]) is int


def test_single_pair_defaults_and_other_alphabets():
    al = load_pkg("aligners")
    assert al.overlap_alignment("ACGT", "TTACGTGG")[3:] == (40, 6)          # end > len(s)
    assert al.overlap_alignment("", "") == ("\nTarget:   \n          \nQuery:    ", "", "", 0, 0)
    # the single-pair kernel compares code points, so any alphabet works like in the reference
    for s, t in [("hello world", "world peace"), ("NNNNACGT", "ACGTNNNN"), ("αβγδ", "γδεζ")]:
        assert al.overlap_alignment(s, t) == orc_any(s, t)


def orc_any(s, t, ma=10, mi=-1, ind=-2 ** 31):
    """Pure-Python restatement of aligners.py:27-82 for non-latin-1 text (tiny inputs only)."""
    n, m = len(s), len(t)
    dp = [[0] * (m + 1) for _ in range(n + 1)]
    tb = [[0] * (m + 1) for _ in range(n + 1)]
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            d = dp[i - 1][j - 1] + (ma if s[i - 1] == t[j - 1] else mi)
            u = dp[i - 1][j] + ind
            l = dp[i][j - 1] + ind
            if d >= u and d >= l:
                dp[i][j], tb[i][j] = d, 0
            elif u >= l:
                dp[i][j], tb[i][j] = u, 1
            else:
                dp[i][j], tb[i][j] = l, 2
    best, bj = -float("inf"), 0
    for j in range(m + 1):
        if dp[n][j] > best:
            best, bj = dp[n][j], j
    i, j, a_s, a_t = n, bj, "", ""
    while i > 0 and j > 0:
        if tb[i][j] == 0:
            a_s, a_t, i, j = s[i - 1] + a_s, t[j - 1] + a_t, i - 1, j - 1
        elif tb[i][j] == 1:
            a_s, a_t, i = s[i - 1] + a_s, "-" + a_t, i - 1
        else:
            a_s, a_t, j = "-" + a_s, t[j - 1] + a_t, j - 1
    return f"\nTarget:   {a_t}\n          {'|' * len(a_t)}\nQuery:    {a_s}", a_s, a_t, int(best), bj


# --------------------------------------------------------------------------- K8 local alignment
def test_local_alignment_golden(golden_local):
    al = load_pkg("aligners")
    for c in golden_local["local"]:
        got = al.local_alignment(c["query"], c["reference"], c["match"], c["mismatch"], c["indel"])
        assert list(got) == c["out"], (c["query"], c["reference"], c["match"], c["mismatch"], c["indel"])
        assert all(type(x) is int for x in got[3:])
    for c in golden_local["wrapped"]:
        got = al.align_read_or_contig_to_reference(c["seq"], c["genome"], c["read_length"])
        assert list(got) == c["out"]


def test_local_alignment_contig_vs_genome_oracle():
    """The evaluation's use (performanceMeasures.py:219): contigs of a few hundred bases against the
    5,386-base genome."""
    al = load_pkg("aligners")
    synth = load_pkg("synth")
    genome = synth.phix_like_genome().tobytes().decode()
    rng = random.Random(3)
    for L in (100, 700, 1800):
        st = rng.randrange(len(genome) - L)
        contig = "".join(ch if rng.random() > 0.02 else rng.choice("ACGT") for ch in genome[st:st + L])
        contig = contig[:L // 2] + contig[L // 2 + 3:]                       # a deletion
        assert al.local_alignment(contig, genome) == orc.local_alignment(contig, genome)


# --------------------------------------------------------------------------- K0 / K1 / K2 / K3
def np_pack(reads, row_words):
    code = {"A": 0, "C": 1, "T": 2, "G": 3}
    out = np.zeros((len(reads), row_words), dtype=np.uint32)
    for u, r in enumerate(reads):
        for i, ch in enumerate(r):
            out[u, i // 16] |= np.uint32(code[ch] << (2 * (i % 16)))
    return out


def np_key(r, k, suffix):
    code = {"A": 0, "C": 1, "T": 2, "G": 3}
    seg = r[-k:] if suffix else r[:k]
    v = 0
    for i, ch in enumerate(seg):
        v |= code[ch] << (2 * i)
    return v


@pytest.mark.parametrize("lo,hi,n", [(0, 40, 500), (1, 150, 700), (990, 1000, 40), (1, 2432, 30)])
def test_pack_reads(eng, lo, hi, n):
    import torch
    rng = random.Random(lo * 1000 + hi)
    reads = rand_reads(rng, n, lo, hi)
    rs, _, offsets = upload(eng, reads)
    eng.check_alphabet(rs)
    got = rs.packed[:rs.n_reads * rs.row_words * 4].view(torch.int32).cpu().numpy().view(np.uint32)
    got = got.reshape(rs.n_reads, rs.row_words)
    assert np.array_equal(got, np_pack(reads, rs.row_words))
    assert np.array_equal(rs.length[:n].cpu().numpy(), np.array([len(r) for r in reads], dtype=np.int32))


def test_pack_rejects_non_acgt(eng, nat):
    for bad in ["ACGTN", "acgt", "ACG-T", "ACGU", "ACGE"]:
        rs, _, _ = upload(eng, ["ACGTACGTACGTACGTACGTA", bad, "TTTT"])
        with pytest.raises(nat.OvlUnsupported):
            eng.check_alphabet(rs)
    g = load_pkg("overlapGraphs")
    with pytest.raises(nat.OvlUnsupported):
        g.construct_overlap_graph_nx_k(["ACGT\u0394", "GT\u0394AC"], k=2)    # beyond one byte per symbol
    # up to four distinct symbols of any kind are re-lettered on the host; more (reads with N, IUPAC codes,
    # mixed case) go through the byte-coded general kernels -- both give the exact graph
    rng = random.Random(12)
    with_n = ["".join(ch if rng.random() > 0.03 else "N" for ch in r) for r in overlapping_reads(rng, 400, 300, 40, 0.01)]
    for reads in (["acgtac", "gtacgg", "acggta"], ["ACGUAC", "GUACGG", "ACGGUA"], ["xyxxy", "xxyxy", "yxyxx", "xyxxy"],
                  ["ACGTN", "GTNAC", "TNACG"], with_n, with_n + ["acgtRYKM", "KMacgtRY"]):
        for k in (2, 5, 0) if len(reads) < 50 else (4, 9):
            G, rc = g.construct_overlap_graph_nx_k(reads, k=k)
            nodes, edges, rc2 = orc.construct_overlap_graph(reads, k)
            assert list(rc.items()) == list(rc2.items()) and list(G.nodes) == nodes
            assert list(G.edges(data=True)) == list(orc.to_networkx(nodes, edges).edges(data=True))


@pytest.mark.parametrize("k", [1, 3, 5, 8, 15, 16, 17, 31, 32])
def test_keys_index_and_join(eng, k):
    import torch
    rng = random.Random(k)
    reads = list(dict.fromkeys(overlapping_reads(rng, 1500, 600, 48, 0.01) + rand_reads(rng, 60, 0, 40) +
                               ["G" * 40, "G" * 33 + "A", "A" + "G" * 35]))
    U = len(reads)
    rs, _, _ = upload(eng, reads)
    idx = eng.kmer_index(rs, k)
    pk = idx.prefix_key[:U].cpu().numpy().view(np.uint64)
    sk = idx.suffix_key[:U].cpu().numpy().view(np.uint64)
    valid = np.array([len(r) >= k for r in reads])
    want_pk = np.array([np_key(r, k, False) if len(r) >= k else 0 for r in reads], dtype=np.uint64)
    want_sk = np.array([np_key(r, k, True) if len(r) >= k else 0 for r in reads], dtype=np.uint64)
    assert np.array_equal(pk[valid], want_pk[valid]) and np.array_equal(sk[valid], want_sk[valid])
    n_idx = int(idx.n_indexed.item())
    assert n_idx == int(valid.sum())
    order = np.argsort(want_pk[valid], kind="stable")
    uids = np.nonzero(valid)[0][order]
    assert np.array_equal(idx.sorted_uid[:n_idx].cpu().numpy().view(np.uint32), uids.astype(np.uint32))
    assert np.array_equal(idx.sorted_key[:n_idx].cpu().numpy().view(np.uint64), want_pk[valid][order])
    pa, pb, first = eng.candidate_pairs(rs, idx, k)
    wa, wb = orc.candidate_pairs(reads, k)
    assert first == 0
    assert np.array_equal(pa.cpu().numpy(), wa) and np.array_equal(pb.cpu().numpy(), wb)
    # sharded: concatenation of the rank slices is the global list
    parts = [eng.candidate_pairs(rs, idx, k, (r, 3)) for r in range(3)]
    assert [p[2] for p in parts] == [len(wa) * r // 3 for r in range(3)]
    assert np.array_equal(torch.cat([p[0] for p in parts]).cpu().numpy(), wa)
    assert np.array_equal(torch.cat([p[1] for p in parts]).cpu().numpy(), wb)


@pytest.mark.parametrize("k", [33, 40, 64, 65, 100])
def test_join_with_hashed_keys_for_long_k(eng, k):
    """k > 32: the index holds 64-bit hashes and the join verifies every match base by base."""
    import torch
    rng = random.Random(k)
    reads = list(dict.fromkeys(overlapping_reads(rng, 1200, 900, 130, 0.004) + rand_reads(rng, 40, 0, 120) +
                               ["G" * 120, "G" * 101 + "A", "A" + "G" * 110]))
    rs, _, _ = upload(eng, reads)
    idx = eng.kmer_index(rs, k)
    pa, pb, _ = eng.candidate_pairs(rs, idx, k)
    wa, wb = orc.candidate_pairs(reads, k)
    assert len(wa) > 0
    assert np.array_equal(pa.cpu().numpy(), wa) and np.array_equal(pb.cpu().numpy(), wb)
    parts = [eng.candidate_pairs(rs, idx, k, (r, 3)) for r in range(3)]
    assert np.array_equal(torch.cat([p[0] for p in parts]).cpu().numpy(), wa)
    assert np.array_equal(torch.cat([p[1] for p in parts]).cpu().numpy(), wb)


def test_all_pairs_k0(eng):
    reads = list(dict.fromkeys(rand_reads(random.Random(1), 40, 1, 12)))
    rs, _, _ = upload(eng, reads)
    pa, pb, _ = eng.candidate_pairs(rs, None, 0)
    wa, wb = orc.candidate_pairs(reads, 0)
    assert np.array_equal(pa.cpu().numpy(), wa) and np.array_equal(pb.cpu().numpy(), wb)


# --------------------------------------------------------------------------- K4 / K5 batch DP
PARAMS = [(10, -1, -2 ** 31), (10, -1, -2), (10, -1, -11), (1, -1, -1), (5, -4, -3), (2, -3, -2),
          (10, -1, 0), (3, 0, -1), (-1, 10, -2), (7, 7, -3), (10, -1, -1000)]


def run_dp(eng, reads, pa, pb, prm, **kw):
    import torch
    rs, bases, offsets = upload(eng, reads)
    ta = torch.from_numpy(pa).to(eng.device)
    tb = torch.from_numpy(pb).to(eng.device)
    score, end = eng.overlap_scores(rs, ta, tb, *prm, **kw)
    ws, we = orc.overlap_pairs(bases, offsets, pa, pb, *prm)
    return score.cpu().numpy(), end.cpu().numpy(), ws, we


@pytest.mark.parametrize("max_len", [1, 7, 25, 38, 64, 100, 150, 151, 300, 1000, 1216, 1217, 2432])
def test_batch_dp_lengths(eng, max_len):
    rng = random.Random(max_len)
    n_reads = 60 if max_len > 300 else 300
    reads = overlapping_reads(rng, n_reads, max(2 * max_len, 50), max_len, 0.03) + rand_reads(rng, 20, 0, max_len)
    reads.append("")
    P = 301 if max_len > 300 else 2001           # odd: exercises the half-empty last couple
    pa = np.array([rng.randrange(len(reads)) for _ in range(P)], dtype=np.int32)
    pb = np.array([rng.randrange(len(reads)) for _ in range(P)], dtype=np.int32)
    for prm in PARAMS[:3] if max_len > 300 else PARAMS:
        s, e, ws, we = run_dp(eng, reads, pa, pb, prm)
        assert np.array_equal(s, ws) and np.array_equal(e, we), (max_len, prm)


def test_batch_dp_every_instantiation(eng, nat):
    """Force every (mode, lanes, columns) kernel that can hold the batch."""
    rng = random.Random(99)
    for max_len, n_pairs in [(24, 1500), (60, 1200), (150, 800), (400, 150), (1000, 40)]:
        reads = overlapping_reads(rng, 120, 3 * max_len, max_len, 0.02) + rand_reads(rng, 10, 0, max_len)
        pa = np.array([rng.randrange(len(reads)) for _ in range(n_pairs)], dtype=np.int32)
        pb = np.array([rng.randrange(len(reads)) for _ in range(n_pairs)], dtype=np.int32)
        tried = 0
        for mode, cols_list in [(1, (19, 25, 32, 38, 76)), (2, (32,))]:
            for lanes in (1, 2, 4, 8, 16, 32):
                for cols in cols_list:
                    if lanes * cols < max(len(r) for r in reads) or (cols == 76 and lanes < 16):
                        continue
                    for prm in [(10, -1, -2 ** 31), (10, -1, -2)]:
                        s, e, ws, we = run_dp(eng, reads, pa, pb, prm, mode=mode, lanes=lanes, cols=cols)
                        assert np.array_equal(s, ws) and np.array_equal(e, we), (max_len, mode, lanes, cols, prm)
                        tried += 1
        assert tried >= 6


def test_batch_dp_long_reads_use_the_anti_diagonal_kernel(eng):
    rng = random.Random(31)
    reads = overlapping_reads(rng, 30, 9000, 4000, 0.02) + rand_reads(rng, 6, 0, 3000) + ["", "ACGT"]
    assert eng.dp_plan(4000)["mode"] == "long-read"
    pa = np.array([rng.randrange(len(reads)) for _ in range(41)], dtype=np.int32)
    pb = np.array([rng.randrange(len(reads)) for _ in range(41)], dtype=np.int32)
    for prm in [(10, -1, -2 ** 31), (10, -1, -2), (2, -3, -2)]:
        s, e, ws, we = run_dp(eng, reads, pa, pb, prm)
        assert np.array_equal(s, ws) and np.array_equal(e, we), prm
    g = load_pkg("overlapGraphs")
    G, rc = g.construct_overlap_graph_nx_k(reads, k=12)
    nodes, edges, rc2 = orc.construct_overlap_graph(reads, 12)
    assert list(G.edges(data=True)) == list(orc.to_networkx(nodes, edges).edges(data=True))


@pytest.mark.parametrize("max_len", [12, 30, 100, 150, 400, 1000, 1500])
def test_byte_coded_dp_equals_oracle(eng, max_len):
    """Byte-coded reads (any alphabet): packed wavefront kernel with XOR/min/IMAD costs when the scores
    fit 16 bits, the anti-diagonal kernel otherwise -- both against the oracle."""
    import torch
    rng = random.Random(1000 + max_len)
    alphabet = "ACGTNRYacgt"
    g = "".join(rng.choice(alphabet) for _ in range(3 * max_len + 20))
    reads = []
    for _ in range(80):
        st = rng.randrange(len(g))
        reads.append("".join(ch if rng.random() > 0.03 else rng.choice(alphabet) for ch in g[st:st + max_len]))
    reads += ["", "N", "ACGTN"]
    bases, offsets = orc.concat_reads(reads)
    rs = eng.upload_reads(bases[:int(offsets[-1])], offsets, code_bits=8)
    n_pairs = 201 if max_len > 400 else 1501
    pa = np.array([rng.randrange(len(reads)) for _ in range(n_pairs)], dtype=np.int32)
    pb = np.array([rng.randrange(len(reads)) for _ in range(n_pairs)], dtype=np.int32)
    ta, tb = torch.from_numpy(pa).to(eng.device), torch.from_numpy(pb).to(eng.device)
    wide = (10, -1000000, -2 ** 31) if max_len <= 400 else (10, -100000, -2 ** 31)     # int32 route, must stay < 2^30
    for prm in [(10, -1, -2 ** 31), (10, -1, -2), (2, -3, -2), (-1, 10, -2), wide]:
        s, e = eng.overlap_scores(rs, ta, tb, *prm)
        ws, we = orc.overlap_pairs(bases, offsets, pa, pb, *prm)
        assert np.array_equal(s.cpu().numpy(), ws) and np.array_equal(e.cpu().numpy(), we), (max_len, prm)


def test_batch_dp_wide_scores_use_int32_kernel(eng):
    rng = random.Random(5)
    reads = overlapping_reads(rng, 100, 400, 120, 0.05)
    pa = np.array([rng.randrange(len(reads)) for _ in range(500)], dtype=np.int32)
    pb = np.array([rng.randrange(len(reads)) for _ in range(500)], dtype=np.int32)
    for prm in [(10, -1000000, -2 ** 31), (1000, -1, -5000), (300, -200, -7), (100000, -100000, -2 ** 31)]:
        assert eng.dp_plan(120, *prm)["mode"] == "int32"
        s, e, ws, we = run_dp(eng, reads, pa, pb, prm)
        assert np.array_equal(s, ws) and np.array_equal(e, we), prm


def test_oversized_pair_list_is_refused_not_oom(eng, nat):
    """k = 0 on 300,000 reads would be 9e10 pairs (~2.6 TiB): a clear error, not an out-of-memory crash."""
    rng = random.Random(8)
    reads = list(dict.fromkeys(rand_reads(rng, 300000, 12, 14)))
    rs, _, _ = upload(eng, reads)
    with pytest.raises(nat.OvlUnsupported, match="GiB"):
        eng.candidate_pairs(rs, None, 0)


def test_batch_dp_unsupported_is_loud(eng, nat):
    reads = ["ACGT" * 10, "CGTA" * 10]
    pa, pb = np.array([0], np.int32), np.array([1], np.int32)
    with pytest.raises(nat.OvlUnsupported):
        run_dp(eng, reads, pa, pb, (2 ** 40, -1, -2))            # would overflow int32 storage
    with pytest.raises(nat.OvlUnsupported):
        upload(eng, ["A" * 20000])                               # longer than any kernel covers


# --------------------------------------------------------------------------- whole builder
def test_graph_builder_golden(golden_graphs, nat):
    g = load_pkg("overlapGraphs")
    for c in golden_graphs:
        G, read_copies = g.construct_overlap_graph_nx_k(c["reads"], k=c["k"])
        assert [[r, n] for r, n in read_copies.items()] == c["read_copies"], c["name"]
        nodes = list(G.nodes)
        assert len(nodes) == c["n_nodes"]
        assert hashlib.sha256("\n".join(nodes).encode()).hexdigest() == c["nodes_sha256"], c["name"]
        idx = {n: i for i, n in enumerate(nodes)}
        got = [[idx[u], idx[v], d["weight"], d["end_position"]] for u, v, d in G.edges(data=True)]
        assert got == c["edges"], c["name"]
        for _, _, d in list(G.edges(data=True))[:5]:
            assert type(d["weight"]) is int and type(d["end_position"]) is int


def test_allpairs_builders_golden(golden_allpairs, capsys):
    """Drop-ins of construct_overlap_graph_string / construct_string_graph vs the live reference's output."""
    g = load_pkg("overlapGraphs")
    for c in golden_allpairs:
        G, rc = g.construct_overlap_graph_string(c["reads"])
        nodes = list(G.nodes)
        assert nodes == c["string_nodes"], c["name"]
        assert [[r, n] for r, n in rc.items()] == c["string_read_copies"]
        idx = {n: i for i, n in enumerate(nodes)}
        assert [[idx[u], idx[v], d["weight"], d["end_position"]] for u, v, d in G.edges(data=True)] == c["string_edges"], c["name"]
        capsys.readouterr()
        H = g.construct_string_graph(c["reads"])
        out = capsys.readouterr().out
        hn = list(H.nodes)
        assert hn == c["sg_nodes"], c["name"]
        hidx = {n: i for i, n in enumerate(hn)}
        assert [[hidx[u], hidx[v], d["weight"], d["end_position"]] for u, v, d in H.edges(data=True)] == c["sg_edges"], c["name"]
        assert [[hidx[p] for p in H.pred[n]] for n in hn] == c["sg_pred"], c["name"]
        assert hashlib.sha256(out.encode()).hexdigest() == c["sg_stdout_sha256"], c["name"]
        for _, _, d in list(H.edges(data=True))[:3]:
            assert type(d["weight"]) is int and type(d["end_position"]) is int


def test_allpairs_builders_vs_oracle_random(eng, capsys):
    g = load_pkg("overlapGraphs")
    rng = random.Random(77)
    reads = overlapping_reads(rng, 120, 80, 20, 0.02) + rand_reads(rng, 10, 1, 8)
    G, rc = g.construct_overlap_graph_string(reads)
    nodes, edges, rc2 = orc.construct_overlap_graph_string(reads)
    G2 = orc.to_networkx(nodes, edges)
    assert list(rc.items()) == list(rc2.items()) and list(G.nodes) == list(G2.nodes)
    assert list(G.edges(data=True)) == list(G2.edges(data=True))
    H = g.construct_string_graph(reads)
    H2 = orc.construct_string_graph(reads)
    capsys.readouterr()
    assert list(H.nodes) == list(H2.nodes)
    assert list(H.edges(data=True)) == list(H2.edges(data=True))
    assert [list(H.pred[n]) for n in H.nodes] == [list(H2.pred[n]) for n in H2.nodes]


def test_batched_read_sets_equal_individual_builds(eng):
    """The sweep as one job (experiments.py:451-539): identical graphs to one build per read set."""
    g = load_pkg("overlapGraphs")
    synth = load_pkg("synth")
    genome = synth.phix_like_genome()
    sets = []
    for i, (n, l, p) in enumerate([(100, 50, 0.001), (316, 100, 0.01), (1000, 150, 0.1), (100, 150, 0.01), (0, 50, 0.0)]):
        b, o = synth.simulate_reads(genome, n, l, p, seed=100 + i) if n else (np.zeros(0, np.uint8), np.zeros(1, np.int64))
        sets.append(synth.to_strings(b, o))
    sets.append(sets[0] + sets[0][:10])                       # a set with duplicate reads
    for k in (5, 10, 15):
        got = g.construct_overlap_graphs_batch(sets, k=k)
        assert len(got) == len(sets)
        for reads, (G, rc) in zip(sets, got):
            G1, rc1 = g.construct_overlap_graph_nx_k(reads, k=k)
            assert list(rc.items()) == list(rc1.items())
            assert list(G.nodes) == list(G1.nodes)
            assert list(G.edges(data=True)) == list(G1.edges(data=True))
            assert [list(G.pred[n]) for n in G.nodes] == [list(G1.pred[n]) for n in G1.nodes]
        nodes, edges, _ = orc.construct_overlap_graph(sets[2], k)
        assert list(got[2][0].edges(data=True)) == list(orc.to_networkx(nodes, edges).edges(data=True))


@pytest.mark.parametrize("reads,k", [
    ([], 5), (["ACGTACGT"], 3), (["ACGTACGT"] * 4, 3), (["ACG", "CG", "A"], 5), (["ACGT", "CGTA"], 8),
    (["ACGTAC", "GTACGT", "ACGTAC", ""], 3), (["", ""], 0), (["A"], 0), (["ACGT", "ACGT", "CGTT"], 0),
    (["AAAAAAAAAA", "AAAAAAAAAA", "AAAAAAAAA"], 2)])
def test_graph_builder_degenerate_inputs(reads, k):
    g = load_pkg("overlapGraphs")
    G, rc = g.construct_overlap_graph_nx_k(reads, k=k)
    nodes, edges, rc2 = orc.construct_overlap_graph(reads, k)
    G2 = orc.to_networkx(nodes, edges)
    assert list(rc.items()) == list(rc2.items())
    assert list(G.nodes) == list(G2.nodes)
    assert list(G.edges(data=True)) == list(G2.edges(data=True))


def test_graph_builder_vs_oracle_with_duplicates(eng):
    g = load_pkg("overlapGraphs")
    rng = random.Random(2024)
    reads = overlapping_reads(rng, 3000, 500, 60, 0.005)       # tiny genome: many duplicate reads
    assert len(set(reads)) < len(reads)
    for k in (4, 9):
        G, rc = g.construct_overlap_graph_nx_k(reads, k=k)
        nodes, edges, rc2 = orc.construct_overlap_graph(reads, k)
        assert list(rc.items()) == list(rc2.items())
        G2 = orc.to_networkx(nodes, edges)
        assert list(G.nodes) == list(G2.nodes)
        assert list(G.edges(data=True)) == list(G2.edges(data=True))
        assert [list(G.pred[n]) for n in G.nodes] == [list(G2.pred[n]) for n in G2.nodes]


def test_edge_rows_insertion_order(eng):
    """Device edge rows come in the reference's insertion order (a, b, copy_a, copy_b)."""
    g = load_pkg("overlapGraphs")
    rng = random.Random(11)
    reads = overlapping_reads(rng, 800, 200, 30, 0.0)
    rc, uniq, counts, edges = g.overlap_edge_rows(reads, k=6)
    nodes, oedges, _ = orc.construct_overlap_graph(reads, 6)
    idx = {n: i for i, n in enumerate(nodes)}
    want = np.array([[idx[u], idx[v], w, e] for u, v, w, e in oedges], dtype=np.int32).reshape(-1, 4)
    assert np.array_equal(edges, want)


def test_fused_epilogue_equals_separate_expansion(eng):
    """K6 fused into the DP epilogue == DP (score/end) followed by the stand-alone expansion."""
    import torch
    synth = load_pkg("synth")
    bases, offsets = synth.simulate_reads(synth.phix_like_genome(), 6000, 100, 0.01, seed=9)
    ub, uo, counts, _ = synth.dedup(bases, offsets)
    assert counts.max() > 1
    rs = eng.upload_reads(ub, uo)
    idx = eng.kmer_index(rs, 5)
    pa, pb, _ = eng.candidate_pairs(rs, idx, 5)
    node_off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=node_off[1:])
    d_copies = torch.from_numpy(counts).to(eng.device)
    d_node = torch.from_numpy(node_off).to(eng.device)
    score, end = eng.overlap_scores(rs, pa, pb)
    for cp, no in [(d_copies, d_node), (None, None)]:
        sep = eng.expand_edges(pa, pb, score, end, cp, no)
        fused = eng.overlap_edges_fused(rs, pa, pb, cp, no)
        assert torch.equal(sep, fused)


def test_sharded_edges_concatenate(eng):
    """Rank slices of the pair list produce edge slices whose concatenation is the 1-GPU list."""
    synth = load_pkg("synth")
    bases, offsets = synth.simulate_reads(synth.phix_like_genome(), 4000, 100, 0.01, seed=5)
    ub, uo, counts, _ = synth.dedup(bases, offsets)
    whole = eng.overlap_edges(ub, uo, counts, k=6)
    parts = [eng.overlap_edges(ub, uo, counts, k=6, shard=(r, 4)) for r in range(4)]
    assert np.array_equal(np.concatenate(parts), whole)


def test_edge_list_hash_matches_host_mirror_and_sees_order(eng):
    import torch
    rng = np.random.default_rng(11)
    rows = rng.integers(-2 ** 31, 2 ** 31 - 1, size=(100_003, 4), dtype=np.int64).astype(np.int32)
    d = torch.from_numpy(rows).to(eng.device)
    h = int(eng.edge_hash(d).item()) & 0xFFFFFFFFFFFFFFFF
    assert h == eng.edge_hash_numpy(rows)
    assert eng.edge_hash_host(rows, chunk_rows=7777) == h
    # shards hashed with their global row offset add up (mod 2^64)
    cut = 41_234
    parts = (eng.edge_hash_numpy(rows[:cut]) + eng.edge_hash_numpy(rows[cut:], first_row=cut)) & 0xFFFFFFFFFFFFFFFF
    assert parts == h
    acc = eng.edge_hash(d[:cut])
    eng.edge_hash(d[cut:], first_row=cut, accum=acc)
    assert int(acc.item()) & 0xFFFFFFFFFFFFFFFF == h
    # a sum of fields cannot see these; the fingerprint must: two rows swapped, one row duplicated over its neighbour
    swapped = rows.copy()
    swapped[[10, 20]] = swapped[[20, 10]]
    assert swapped.astype(np.int64).sum() == rows.astype(np.int64).sum()
    assert int(eng.edge_hash(torch.from_numpy(swapped).to(eng.device)).item()) & 0xFFFFFFFFFFFFFFFF != h
    comp = rows.copy()
    comp[5, 2] += 1
    comp[6, 2] -= 1                                     # compensating errors
    assert int(eng.edge_hash(torch.from_numpy(comp).to(eng.device)).item()) & 0xFFFFFFFFFFFFFFFF != h
    assert int(eng.edge_hash(d[:0]).item()) == 0


# --------------------------------------------------------------------------- K0-K3 as one job
def _job(eng, reads, k, counts=None, shard=(0, 1), segments=None, n_segments=1):
    import torch
    bases, offsets = orc.concat_reads(reads)
    ascii_dev, off_dev, U, max_len = eng._upload_ascii(bases[:int(offsets[-1])] if len(reads) else bases[:0], offsets)
    copies = node_off = None
    if counts is not None and max(counts) > 1:
        c = np.asarray(counts, dtype=np.int32)
        no = np.zeros(len(c) + 1, np.int64)
        np.cumsum(c, out=no[1:])
        copies, node_off = torch.from_numpy(c).to(eng.device), torch.from_numpy(no).to(eng.device)
    seg = torch.from_numpy(np.asarray(segments, dtype=np.int32)).to(eng.device) if segments is not None else None
    return eng.build_candidates(ascii_dev, off_dev, U, max_len, k, copies, node_off, shard, seg, n_segments)


@pytest.mark.parametrize("k", [1, 2, 5, 8, 10, 11, 12, 15, 16, 31, 32])
def test_one_call_job_keys_index_table_and_pairs(eng, k):
    """ovl_candidates_build (fused pack + keys, wide-digit sort, bucket table, table join) against NumPy / the oracle."""
    import torch
    rng = random.Random(100 + k)
    reads = list(dict.fromkeys(overlapping_reads(rng, 2500, 700, 70, 0.01) + rand_reads(rng, 200, 0, 140) +
                               ["G" * 40, "G" * 33 + "A", "A" + "G" * 35, "C" * 64, "C" * 65, "T" * 63 + "A", "A" * 96 + "C" * 32]))
    U = len(reads)
    cand = _job(eng, reads, k)
    idx = cand.index
    pk = idx.prefix_key[:U].cpu().numpy().view(np.uint64)
    sk = idx.suffix_key[:U].cpu().numpy().view(np.uint64)
    valid = np.array([len(r) >= k for r in reads])
    want_pk = np.array([np_key(r, k, False) if len(r) >= k else 0 for r in reads], dtype=np.uint64)
    want_sk = np.array([np_key(r, k, True) if len(r) >= k else 0 for r in reads], dtype=np.uint64)
    assert np.array_equal(pk[valid], want_pk[valid]) and np.array_equal(sk[valid], want_sk[valid])
    got_rows = cand.rs.packed[:U * cand.rs.row_words * 4].view(torch.int32).cpu().numpy().view(np.uint32).reshape(U, -1)
    assert np.array_equal(got_rows, np_pack(reads, cand.rs.row_words))
    n_idx = int(idx.n_indexed.item())
    assert n_idx == int(valid.sum())
    order = np.argsort(want_pk[valid], kind="stable")
    uids = np.nonzero(valid)[0][order]
    skeys = want_pk[valid][order]
    assert np.array_equal(idx.sorted_uid[:n_idx].cpu().numpy().view(np.uint32), uids.astype(np.uint32))
    assert np.array_equal(idx.sorted_key[:n_idx].cpu().numpy().view(np.uint64), skeys)
    assert idx.pos_of is None            # the one-call job does not build it (the granular ovl_index_build can)
    # the bucket table: first sorted position whose key prefix is >= t
    shift = idx.key_bits - idx.table_bits
    table = idx.table.cpu().numpy()
    want_table = np.searchsorted(skeys >> np.uint64(shift), np.arange((1 << idx.table_bits) + 1, dtype=np.uint64), side="left")
    assert np.array_equal(table, want_table.astype(np.int32))
    wa, wb = orc.candidate_pairs(reads, k)
    assert cand.total_pairs == len(wa) == cand.total_edges and (cand.p_begin, cand.p_end) == (0, len(wa))
    pa, pb = eng.fill_pairs(cand)
    assert np.array_equal(pa.cpu().numpy(), wa) and np.array_equal(pb.cpu().numpy(), wb)
    # the 65 cuts: monotone, from p_begin to p_end, equal parts
    assert cand.cut_pairs[0] == 0 and cand.cut_pairs[-1] == len(wa) and cand.cut_pairs == sorted(cand.cut_pairs)
    assert cand.cut_pairs[32] == len(wa) // 64 * 32 + len(wa) % 64 * 32 // 64 and cand.cut_edges == cand.cut_pairs
    # sharded jobs tile the list
    parts = [_job(eng, reads, k, shard=(r, 3)) for r in range(3)]
    assert [(c.p_begin, c.p_end) for c in parts] == [(len(wa) * r // 3, len(wa) * (r + 1) // 3) for r in range(3)]
    filled = [eng.fill_pairs(c) for c in parts]
    assert np.array_equal(torch.cat([f[0] for f in filled]).cpu().numpy(), wa)
    assert np.array_equal(torch.cat([f[1] for f in filled]).cpu().numpy(), wb)
    # the granular entry points (separate pack, keys, index, join calls) give the same list
    rs, _, _ = upload(eng, reads)
    gi = eng.kmer_index(rs, k)
    ga, gb, _ = eng.candidate_pairs(rs, gi, k)
    assert torch.equal(ga, pa) and torch.equal(gb, pb)
    # every indexed read knows its own sorted position
    pos_of = gi.pos_of[:U].cpu().numpy()
    assert np.array_equal(pos_of[uids], np.arange(n_idx))


def test_one_call_job_suffix_key_straddles_groups_and_warps(eng):
    """The fused suffix key: k-mers that straddle two 64-base groups, incl. groups packed by different warps."""
    rng = random.Random(77)
    reads = []
    for L in list(range(1, 200)) + [255, 256, 257, 319, 320, 321, 1000, 2047, 2048, 2049]:
        reads.append("".join(rng.choice("ACGT") for _ in range(L)))
    reads = list(dict.fromkeys(reads))
    U = len(reads)
    for k in (1, 7, 17, 31, 32):
        cand = _job(eng, reads, k)
        sk = cand.index.suffix_key[:U].cpu().numpy().view(np.uint64)
        pk = cand.index.prefix_key[:U].cpu().numpy().view(np.uint64)
        for u, r in enumerate(reads):
            if len(r) >= k:
                assert int(sk[u]) == np_key(r, k, True) and int(pk[u]) == np_key(r, k, False), (k, len(r))
    # 600 reads of 65..96 bases: row = 2 groups, so every 16th read's second group starts a warp
    reads = list(dict.fromkeys("".join(rng.choice("ACGT") for _ in range(rng.randint(65, 96))) for _ in range(600)))
    for k in (20, 32):
        cand = _job(eng, reads, k)
        sk = cand.index.suffix_key[:len(reads)].cpu().numpy().view(np.uint64)
        assert [int(x) for x in sk] == [np_key(r, k, True) for r in reads]


def test_one_call_job_edge_offsets_with_copies(eng):
    """Reads with copies: the implicit edge offsets of the join index (no per-pair offset array) reproduce the
    explicit scan of copies[a] * copies[b], for the whole list, for shards and for D2H chunks."""
    import torch
    rng = random.Random(5)
    base = overlapping_reads(rng, 1500, 500, 50, 0.01) + ["ACGTACGTAC", "CGTACGTACG", "GTACG"]
    base = list(dict.fromkeys(base))
    counts = [rng.choice([1, 1, 1, 2, 3, 7]) for _ in base]
    reads = [r for r, c in zip(base, counts) for _ in range(c)]
    rng.shuffle(reads)
    rc = orc.dedup_reads(reads)
    uniq, cnts = list(rc.keys()), list(rc.values())
    for k in (3, 6, 12):
        nodes, edges, _ = orc.construct_overlap_graph(reads, k)
        nid = {n: i for i, n in enumerate(nodes)}
        want = np.array([[nid[u], nid[v], w, e] for u, v, w, e in edges], dtype=np.int32).reshape(-1, 4)
        cand = _job(eng, uniq, k, cnts)
        assert cand.total_edges == len(want) and cand.edge_base is not None
        pa, pb = eng.fill_pairs(cand)
        got = eng.candidate_edges(cand, pa, pb)
        assert np.array_equal(got.cpu().numpy(), want)
        # cut_edges are the true edge offsets of cut_pairs
        c = np.asarray(cnts, dtype=np.int64)
        per_pair = c[pa.cpu().numpy()] * c[pb.cpu().numpy()]
        off = np.concatenate([[0], np.cumsum(per_pair)])
        assert cand.cut_edges == [int(off[p]) for p in cand.cut_pairs]
        # shards: rows land at their global offsets
        parts = []
        for r in range(4):
            cs = _job(eng, uniq, k, cnts, shard=(r, 4))
            sa, sb = eng.fill_pairs(cs)
            rows = eng.candidate_edges(cs, sa, sb)
            assert cs.e_begin == int(off[cs.p_begin]) and cs.e_end == int(off[cs.p_end])
            parts.append(rows)
        assert np.array_equal(torch.cat(parts).cpu().numpy(), want)
        # host path (chunked D2H) and the public host call
        host = eng.candidate_edges_to_host(cand, pa, pb, min_chunked_pairs=0)                 # 7 shrinking chunks
        assert np.array_equal(host, want)
        host = eng.candidate_edges_to_host(cand, pa, pb, chunk_pairs=max(1, len(pa) // 5), min_chunked_pairs=0)
        assert np.array_equal(host, want)
        b, o = orc.concat_reads(uniq)
        assert np.array_equal(eng.overlap_edges(b, o, np.asarray(cnts, np.int32), k), want)


def test_one_call_job_finite_indel_and_all_a_genome(eng):
    """Degenerate buckets: a homopolymer genome puts every read in ONE bucket (every read is its own candidate's
    neighbour); results must still equal the oracle."""
    reads = ["A" * n for n in range(3, 60)] + ["A" * 20 + "C", "C" + "A" * 20]
    counts = [1 + (i % 3) for i in range(len(reads))]
    many = [r for r, c in zip(reads, counts) for _ in range(c)]
    for k in (3, 5):
        nodes, edges, _ = orc.construct_overlap_graph(many, k)
        nid = {n: i for i, n in enumerate(nodes)}
        want = np.array([[nid[u], nid[v], w, e] for u, v, w, e in edges], dtype=np.int32).reshape(-1, 4)
        b, o = orc.concat_reads(reads)
        got = eng.overlap_edges(b, o, np.asarray(counts, np.int32), k)
        assert np.array_equal(got, want)


def test_local_alignment_batch_golden_and_oracle(golden_local):
    """K8 batch (one CTA per query, one launch): every golden case through the batch entry points, and a contig set
    against the 5,386-base genome the way performanceMeasures.py:219-221 uses it -- identical to per-call results."""
    al = load_pkg("aligners")
    groups = {}
    for c in golden_local["local"]:
        groups.setdefault((c["reference"], c["match"], c["mismatch"], c["indel"]), []).append(c)
    n_batched = 0
    for (ref, ma, mi, ind), cases in groups.items():
        got = al.local_alignment_batch([c["query"] for c in cases], ref, ma, mi, ind)
        for c, g in zip(cases, got):
            assert list(g) == c["out"], (c["query"], ref, ma, mi, ind)
            assert all(type(x) is int for x in g[3:])
        n_batched += len(cases)
    assert n_batched == len(golden_local["local"])
    wgroups = {}
    for c in golden_local["wrapped"]:
        wgroups.setdefault((c["genome"], c["read_length"]), []).append(c)
    for (genome, rl), cases in wgroups.items():
        got = al.align_reads_or_contigs_to_reference([c["seq"] for c in cases], genome, rl)
        for c, g in zip(cases, got):
            assert list(g) == c["out"]
    # contigs vs genome: lengths 1..1024 in the batch kernel, one longer contig through the per-call path, an empty one
    synth = load_pkg("synth")
    genome = synth.phix_like_genome().tobytes().decode()
    rng = random.Random(9)
    contigs = [""]
    for L in [1, 5, 31, 32, 33, 60, 99, 100, 101, 150, 257, 511, 700, 1023, 1024, 1500] + [rng.randint(20, 400) for _ in range(40)]:
        st = rng.randrange(len(genome) - L)
        c = "".join(ch if rng.random() > 0.03 else rng.choice("ACGT") for ch in genome[st:st + L])
        if L > 40:
            c = c[:L // 2] + c[L // 2 + 2:]                                  # a deletion
        contigs.append(c)
    contigs.append(genome[-80:])                                             # at the genome's end
    contigs.append(contigs[5])                                               # a repeated contig
    read_length = 100
    got = al.align_reads_or_contigs_to_reference(contigs, genome, read_length)
    for c, g in zip(contigs, got):
        n = len(c)
        if n < read_length:                                                  # aligners.py:191-199
            w = orc.local_alignment(c, genome[-n:]) if n else orc.local_alignment(c, genome[len(genome):])
            want = (w[0], w[1], w[2], w[3], len(genome) - n + w[4], len(genome) - n + w[5])
        else:
            want = orc.local_alignment(c, genome)
        assert tuple(g) == tuple(want), n
        assert tuple(al.align_read_or_contig_to_reference(c, genome, read_length)) == tuple(want)
    # the prefetch cache serves the reference's own per-contig loop
    assert al.prefetch_alignments(contigs, genome, read_length) == len(set(contigs))
    eng = load_pkg("engine").get_engine()
    before = eng.launches
    for c, g in zip(contigs, got):
        assert tuple(al.align_read_or_contig_to_reference(c, genome, read_length)) == tuple(g)
    assert eng.launches == before                                            # no kernel ran
    al._PREFETCHED.clear()


@pytest.mark.parametrize("n,l,p,G", [(1, 1, 0.0, 1), (500, 50, 0.001, 5386), (3162, 150, 0.1, 5386), (20000, 100, 0.01, 123457),
                                     (40, 1000, 0.02, 900)])
def test_device_read_simulator_equals_numpy_mirror(eng, n, l, p, G):
    """ovl_simulate_reads (generateErrorFreeReads.py:22-52 + generateErrorProneReads.py:4-45 on a counter-based stream)
    produces the bytes of synth.simulate_reads_counter, and they are what the builder consumes."""
    synth = load_pkg("synth")
    genome = synth.random_genome(G, 5)
    for seed in (0, 12345, 2 ** 63 + 11):
        want_b, want_o = synth.simulate_reads_counter(genome, n, l, p, seed)
        ascii_dev, off_dev = eng.simulate_reads(genome, n, l, p, seed)
        got_o = off_dev.cpu().numpy()
        assert np.array_equal(got_o, want_o)
        assert np.array_equal(ascii_dev[:int(got_o[-1])].cpu().numpy(), want_b)
    lens = want_o[1:] - want_o[:-1]
    assert lens.min() >= 1 and lens.max() <= l                       # end-truncated, never empty (:42)
    if n >= 500:
        clean_b, _ = synth.simulate_reads_counter(genome, n, l, 0.0, 2 ** 63 + 11)
        rate = float((clean_b != want_b).mean())
        assert abs(rate - p) < 0.25 * p + 0.002                      # substitution rate ~ p, always to another base
    # device reads straight into the one-call job == the host path on the same reads
    if n >= 500 and l <= 150:
        reads = synth.to_strings(want_b, want_o)
        ub, uo, counts, _ = synth.dedup(want_b, want_o)
        if counts.max() == 1:                                        # the device job has no de-duplication of its own
            cand = eng.build_candidates(ascii_dev, off_dev, n, l, 5)
            pa, pb = eng.fill_pairs(cand)
            wa, wb = orc.candidate_pairs(reads, 5)
            assert np.array_equal(pa.cpu().numpy(), wa) and np.array_equal(pb.cpu().numpy(), wb)


def test_cycle_removal_with_device_prepass_golden(golden_graphs, golden_cycles, eng):
    """overlapGraphs.remove_cycles_from_graph (device sink-peeling pre-pass + find_cycle on the survivors) removes the
    edges the live reference removed, in the same order, on every golden graph -- and leaves the same DAG."""
    import networkx as nx
    g = load_pkg("overlapGraphs")
    by_name = {c["name"]: c for c in golden_graphs}
    checked = 0
    for cyc in golden_cycles:
        c = by_name[cyc["name"]]
        G, _ = g.construct_overlap_graph_nx_k(c["reads"], k=c["k"])
        idx = {v: i for i, v in enumerate(G.nodes)}
        removed = []
        orig = G.remove_edge
        G.remove_edge = lambda u, v, _o=orig, _r=removed, _i=idx: (_r.append([_i[u], _i[v]]), _o(u, v))[1]
        out = g.remove_cycles_from_graph(G)
        del G.remove_edge
        assert out is G
        assert removed == cyc["removed"], cyc["name"]
        assert G.number_of_edges() == cyc["edges_left"] and nx.is_directed_acyclic_graph(G)
        checked += len(removed)
    assert checked > 1500
    # the pre-pass itself: survivors == nodes that can reach a cycle (brute force with NetworkX)
    rng = random.Random(2)
    for n, m in ((1, 0), (30, 40), (200, 260), (400, 1200)):
        Gd = nx.gnm_random_graph(n, m, seed=n, directed=True)
        src = np.array([u for u, v in Gd.edges], dtype=np.int32)
        dst = np.array([v for u, v in Gd.edges], dtype=np.int32)
        keep, rounds = eng.trim_sinks(src, dst, n)
        on_cycle = set()
        for comp in nx.strongly_connected_components(Gd):
            if len(comp) > 1:
                on_cycle |= comp
        reach = set(on_cycle)
        for v in on_cycle:
            reach |= nx.ancestors(Gd, v)
        assert set(np.nonzero(keep)[0].tolist()) == reach, (n, m)
