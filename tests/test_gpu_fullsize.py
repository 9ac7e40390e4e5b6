"""Full-size checks at BASELINE.json configs[1] (PhiX-like, N=50,000, l=150, p=0.01) and a slice of
configs[3] (l=1000): size-independent properties plus oracle parity on random samples."""
import numpy as np
import pytest

from conftest import load_pkg, has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]

from oracle import overlap_oracle as orc  # noqa: E402  (the checker)


@pytest.fixture(scope="module")
def phix50k():
    import torch
    synth = load_pkg("synth")
    eng = load_pkg("engine").get_engine()
    bases, offsets = synth.make_workload("phix_n50000_l150")
    ub, uo, counts, _ = synth.dedup(bases, offsets)
    rs = eng.upload_reads(ub, uo)
    idx = eng.kmer_index(rs, 5)
    pa, pb, _ = eng.candidate_pairs(rs, idx, 5)
    score, end = eng.overlap_scores(rs, pa, pb)
    torch.cuda.synchronize()
    return dict(eng=eng, ub=ub, uo=uo, counts=counts, rs=rs, idx=idx, pa=pa, pb=pb, score=score, end=end)


def test_candidate_list_properties(phix50k):
    d = phix50k
    pa, pb = d["pa"].cpu().numpy().astype(np.int64), d["pb"].cpu().numpy().astype(np.int64)
    U = len(d["counts"])
    # ordered by (a, b), no self pairs, no duplicates
    key = pa * U + pb
    assert np.all(np.diff(key) > 0)
    assert np.all(pa != pb)
    # every pair satisfies suffix_k(a) == prefix_k(b), and the count equals an independent NumPy join
    k = 5
    lens = (d["uo"][1:] - d["uo"][:-1])
    assert np.all(lens[pa] >= k) and np.all(lens[pb] >= k)
    suf = np.stack([d["ub"][d["uo"][pa + 1] - k + i] for i in range(k)], axis=1)
    pre = np.stack([d["ub"][d["uo"][pb] + i] for i in range(k)], axis=1)
    assert np.array_equal(suf, pre)
    valid = np.nonzero(lens >= k)[0]
    code = np.zeros(256, np.int64)
    code[np.frombuffer(b"ACGT", np.uint8)] = np.arange(4)
    pk = sum(code[d["ub"][d["uo"][valid] + i]] * 4 ** i for i in range(k))
    sk = sum(code[d["ub"][d["uo"][valid + 1] - k + i]] * 4 ** i for i in range(k))
    hist = np.bincount(pk, minlength=4 ** k)
    expect = int(hist[sk].sum() - (pk == sk).sum())
    assert len(pa) == expect


def test_scores_sample_vs_oracle_and_invariants(phix50k):
    d = phix50k
    score, end = d["score"].cpu().numpy(), d["end"].cpu().numpy()
    pa, pb = d["pa"].cpu().numpy(), d["pb"].cpu().numpy()
    lens = (d["uo"][1:] - d["uo"][:-1])
    # default scoring: the k shared bases alone give >= 10*k - ... ; bounds from the recurrence
    assert score.min() >= 0 and np.all(score <= 10 * np.minimum(lens[pa], lens[pb]))
    assert np.all(end >= 0) and np.all(end <= lens[pb])
    assert np.all(score >= 10 * 5 - 0)          # the shared 5-mer is an exact overlap of length 5
    rng = np.random.default_rng(3)
    sel = rng.choice(len(pa), size=20000, replace=False)
    ws, we = orc.overlap_pairs(d["ub"], d["uo"], pa[sel], pb[sel])
    assert np.array_equal(score[sel], ws) and np.array_equal(end[sel], we)
    # finite indel on the same sample (gaps can only raise the score)
    import torch
    eng = d["eng"]
    ta, tb = torch.from_numpy(pa[sel]).to(eng.device), torch.from_numpy(pb[sel]).to(eng.device)
    s2, e2 = eng.overlap_scores(d["rs"], ta, tb, 10, -1, -2)
    ws2, we2 = orc.overlap_pairs(d["ub"], d["uo"], pa[sel], pb[sel], 10, -1, -2)
    assert np.array_equal(s2.cpu().numpy(), ws2) and np.array_equal(e2.cpu().numpy(), we2)
    assert np.all(ws2 >= ws)


def test_edge_list_checksum_is_shard_invariant(phix50k):
    d = phix50k
    eng = d["eng"]
    whole = eng.overlap_edges(d["ub"], d["uo"], d["counts"], k=5)
    E = int((d["counts"][d["pa"].cpu().numpy()].astype(np.int64) * d["counts"][d["pb"].cpu().numpy()]).sum())
    assert whole.shape == (E, 4)
    chk = int(whole.astype(np.int64).sum())
    parts = [eng.overlap_edges(d["ub"], d["uo"], d["counts"], k=5, shard=(r, 8)) for r in range(8)]
    assert sum(int(p.astype(np.int64).sum()) for p in parts) == chk
    assert np.array_equal(np.concatenate(parts), whole)
    # idempotence: a second run gives the same bytes
    assert np.array_equal(eng.overlap_edges(d["ub"], d["uo"], d["counts"], k=5), whole)


def test_long_reads_sample_vs_oracle():
    """configs[3] shape: l = 1000, p = 0.02 (warp-per-couple instantiation), smaller genome."""
    import torch
    synth = load_pkg("synth")
    eng = load_pkg("engine").get_engine()
    genome = synth.random_genome(200_000, 11)
    bases, offsets = synth.simulate_reads(genome, 8000, 1000, 0.02, seed=4)
    ub, uo, counts, _ = synth.dedup(bases, offsets)
    rs = eng.upload_reads(ub, uo)
    assert eng.dp_plan(rs.max_len)["lanes"] == 32
    idx = eng.kmer_index(rs, 8)
    pa, pb, _ = eng.candidate_pairs(rs, idx, 8)
    score, end = eng.overlap_scores(rs, pa, pb)
    n = min(int(pa.shape[0]), 1500)
    sel = np.random.default_rng(1).choice(int(pa.shape[0]), size=n, replace=False)
    pa_h, pb_h = pa.cpu().numpy()[sel], pb.cpu().numpy()[sel]
    ws, we = orc.overlap_pairs(ub, uo, pa_h, pb_h)
    assert np.array_equal(score.cpu().numpy()[sel], ws) and np.array_equal(end.cpu().numpy()[sel], we)


def test_config2_one_million_reads_properties():
    """BASELINE.json configs[2] at full size (1 M reads x 150 bp, k = 5): size-independent properties of the
    index / join on all 9.3e8 pairs, and oracle parity of the DP on random samples of them."""
    import torch
    synth = load_pkg("synth")
    eng = load_pkg("engine").get_engine()
    bases, offsets = synth.make_workload("ecoli_n1m_l150")
    ub, uo, counts, _ = synth.dedup(bases, offsets)
    U = len(counts)
    rs = eng.upload_reads(ub, uo)
    eng.check_alphabet(rs)
    k = 5
    idx = eng.kmer_index(rs, k)
    pa, pb, _ = eng.candidate_pairs(rs, idx, k)
    P = int(pa.shape[0])
    # (1) the pair count equals an independent NumPy join on the host
    lens = (uo[1:] - uo[:-1])
    valid = np.nonzero(lens >= k)[0]
    code = np.zeros(256, np.int64)
    code[np.frombuffer(b"ACGT", np.uint8)] = np.arange(4)
    pk = sum(code[ub[uo[valid] + i]] * 4 ** i for i in range(k))
    sk = sum(code[ub[uo[valid + 1] - k + i]] * 4 ** i for i in range(k))
    hist = np.bincount(pk, minlength=4 ** k)
    assert P == int(hist[sk].sum() - (pk == sk).sum())
    # (2) ordered by (a, b), no self pairs -- checked on the device over all pairs
    a64, b64 = pa.to(torch.int64), pb.to(torch.int64)
    key = a64 * U + b64
    assert bool((key[1:] > key[:-1]).all())
    assert bool((pa != pb).all())
    del key, a64, b64
    # (3) suffix_k(a) == prefix_k(b) for every pair, via the device keys
    assert bool((idx.suffix_key[:U][pa.long()] == idx.prefix_key[:U][pb.long()]).all())
    # (4) sharded slices tile the list exactly
    q = [eng.candidate_pairs(rs, idx, k, (r, 8)) for r in (0, 3, 7)]
    assert q[0][2] == 0 and q[1][2] == P * 3 // 8 and q[2][2] == P * 7 // 8
    assert torch.equal(q[1][0], pa[P * 3 // 8:P * 4 // 8]) and torch.equal(q[2][1], pb[P * 7 // 8:])
    del q
    # (5) DP parity on random samples of the real pair list, default and finite indel
    g = torch.Generator(device="cpu").manual_seed(7)
    sel = torch.randint(0, P, (6000,), generator=g).to(eng.device)
    sa, sb = pa[sel].contiguous(), pb[sel].contiguous()
    for prm in [(10, -1, -2 ** 31), (10, -1, -2)]:
        s, e = eng.overlap_scores(rs, sa, sb, *prm)
        ws, we = orc.overlap_pairs(ub, uo, sa.cpu().numpy(), sb.cpu().numpy(), *prm)
        assert np.array_equal(s.cpu().numpy(), ws) and np.array_equal(e.cpu().numpy(), we)
    # (6) score bounds over a 20 M-pair slice: 10*k <= score <= 10*min(len), 0 <= end <= len(b)
    n = min(P, 20_000_000)
    s, e = eng.overlap_scores(rs, pa[:n].contiguous(), pb[:n].contiguous())
    ld = rs.length[:U]
    la, lb = ld[pa[:n].long()], ld[pb[:n].long()]
    assert bool((s >= 10 * k).all()) and bool((s <= 10 * torch.minimum(la, lb)).all())
    assert bool((e >= 0).all()) and bool((e <= lb).all())


def test_config3_long_reads_full_size():
    """BASELINE.json configs[3] at full size (200,000 reads x 1,000 bp, p = 0.02, 4.6 Mb genome): the pair lists
    at k = 5 (3.9e7 pairs) and k = 8 against an independent NumPy join, their order, and oracle parity of the
    32-lane DP instantiation on random samples (default and finite indel) plus the whole k = 8 list's bounds."""
    import torch
    synth = load_pkg("synth")
    eng = load_pkg("engine").get_engine()
    bases, offsets = synth.make_workload("ecoli_n200k_l1000")
    ub, uo, counts, _ = synth.dedup(bases, offsets)
    U = len(counts)
    rs = eng.upload_reads(ub, uo)
    eng.check_alphabet(rs)
    assert rs.max_len == 1000 and eng.dp_plan(rs.max_len) == {"mode": "packed16", "lanes": 32, "cols": 32}
    lens = (uo[1:] - uo[:-1])
    code = np.zeros(256, np.int64)
    code[np.frombuffer(b"ACGT", np.uint8)] = np.arange(4)
    for k, n_oracle in ((5, 2000), (8, 2000)):
        idx = eng.kmer_index(rs, k)
        pa, pb, _ = eng.candidate_pairs(rs, idx, k)
        P = int(pa.shape[0])
        # (1) pair count == independent host join
        valid = np.nonzero(lens >= k)[0]
        pk = sum(code[ub[uo[valid] + i]] * 4 ** i for i in range(k))
        sk = sum(code[ub[uo[valid + 1] - k + i]] * 4 ** i for i in range(k))
        hist = np.bincount(pk, minlength=4 ** k)
        assert P == int(hist[sk].sum() - (pk == sk).sum())
        assert P > (30_000_000 if k == 5 else 400_000)
        # (2) ordered by (a, b), no self pairs; (3) key equality for every pair
        key = pa.to(torch.int64) * U + pb.to(torch.int64)
        assert bool((key[1:] > key[:-1]).all()) and bool((pa != pb).all())
        del key
        assert bool((idx.suffix_key[:U][pa.long()] == idx.prefix_key[:U][pb.long()]).all())
        # (4) DP parity on random samples of the real list: default scoring and a finite indel
        g = torch.Generator(device="cpu").manual_seed(100 + k)
        sel = torch.randint(0, P, (n_oracle,), generator=g).to(eng.device)
        sa, sb = pa[sel].contiguous(), pb[sel].contiguous()
        for prm in [(10, -1, -2 ** 31), (10, -1, -2)]:
            s, e = eng.overlap_scores(rs, sa, sb, *prm)
            ws, we = orc.overlap_pairs(ub, uo, sa.cpu().numpy(), sb.cpu().numpy(), *prm)
            assert np.array_equal(s.cpu().numpy(), ws) and np.array_equal(e.cpu().numpy(), we)
        if k == 8:
            # (5) the whole list: score / end bounds, and the fused edge rows equal (a, b, score, end)
            s, e = eng.overlap_scores(rs, pa, pb)
            ld = rs.length[:U]
            la, lb = ld[pa.long()], ld[pb.long()]
            assert bool((s >= 10 * k).all()) and bool((s <= 10 * torch.minimum(la, lb)).all())
            assert bool((e >= 0).all()) and bool((e <= lb).all())
            rows = eng.overlap_edges_fused(rs, pa, pb)
            assert torch.equal(rows, torch.stack((pa, pb, s, e), dim=1))
        del pa, pb, idx


@pytest.mark.parametrize("k", [5, 6])
def test_index_beyond_one_scan_tile_of_sort_ctas(k):
    """17 M short reads: more sort CTAs (4,151 of 4,096 elements) than one tile of the per-digit scan holds, with a
    single-pass index whose bucket table comes straight from the digit totals (k = 5) and a two-pass one (k = 6).
    Checked against torch's stable sort of the same keys: sorted keys, uid order inside equal keys, bucket table."""
    import torch
    eng = load_pkg("engine").get_engine()
    dev = eng.device
    U, L = 17_000_000, 12
    g = torch.Generator(device=dev).manual_seed(99 + k)
    letters = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ascii_dev = torch.empty(U * L + 64, dtype=torch.uint8, device=dev)
    ascii_dev[:U * L] = letters[torch.randint(0, 4, (U * L,), device=dev, generator=g)]
    off_dev = torch.arange(U + 1, dtype=torch.int64, device=dev) * L
    rs = eng.pack_reads(ascii_dev, off_dev, U, L)
    idx = eng.kmer_index(rs, k)
    n = int(idx.n_indexed.item())
    assert n == U
    want_key, want_uid = torch.sort(idx.prefix_key[:U], stable=True)          # keys < 2^12: the int64 view orders like uint64
    assert torch.equal(idx.sorted_key[:U], want_key)
    assert torch.equal(idx.sorted_uid[:U].to(torch.int64), want_uid)
    assert idx.table is not None and idx.table_bits == 2 * k
    t = torch.arange((1 << idx.table_bits) + 1, device=dev, dtype=torch.int64)
    assert torch.equal(idx.table.to(torch.int64), torch.searchsorted(want_key, t))
