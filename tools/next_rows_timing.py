"""Timings of the 'next rows'  on the GPU box: local_alignment and the batched sweep."""
import importlib, os, sys, time, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "genome-assembly-using-overlap-graphs_b200"
import numpy as np
import torch
synth = importlib.import_module(PKG + ".synth")
al = importlib.import_module(PKG + ".aligners")
og = importlib.import_module(PKG + ".overlapGraphs")
from oracle import overlap_oracle as orc

genome = synth.phix_like_genome().tobytes().decode()
for L in (100, 1000, 5000):
    contig = genome[200:200 + L]
    al.local_alignment(contig, genome)                       # warm
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        out = al.local_alignment(contig, genome)
    torch.cuda.synchronize(); t_gpu = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter(); ref = orc.local_alignment(contig, genome); t_cpu = time.perf_counter() - t0
    assert out == ref
    print(f"local_alignment {L} x {len(genome)}: gpu {t_gpu*1e3:.2f} ms ({L*len(genome)/t_gpu/1e6:.0f} MCUPS), oracle C 1 thread {t_cpu*1e3:.1f} ms")

# the graph-build step of the reference's sweep grid (experiments.py:49-53), 2 iterations
sets = []
g = synth.phix_like_genome()
i = 0
for n in (100, 316, 1000, 3162, 10000):
    for l in (50, 100, 150):
        for p in (0.001, 0.01, 0.1):
            for it in range(2):
                b, o = synth.simulate_reads(g, n, l, p, seed=1000 + i); i += 1
                sets.append(synth.to_strings(b, o))
print(len(sets), "read sets,", sum(len(s) for s in sets), "reads")
for k in (5, 10, 15):
    og.construct_overlap_graphs_batch(sets[:4], k=k)         # warm
    torch.cuda.synchronize(); t0 = time.perf_counter()
    one = [og.construct_overlap_graph_nx_k(s, k=k) for s in sets]
    torch.cuda.synchronize(); t_one = time.perf_counter() - t0
    t0 = time.perf_counter()
    bat = og.construct_overlap_graphs_batch(sets, k=k)
    torch.cuda.synchronize(); t_bat = time.perf_counter() - t0
    assert all(list(a[0].edges(data=True)) == list(b[0].edges(data=True)) for a, b in zip(one, bat))
    print(f"k={k}: {len(sets)} graph builds one by one {t_one:.2f} s, as one batch {t_bat:.2f} s, edges {sum(x[0].number_of_edges() for x in bat)}")

# general kernels: long reads (> 2,432 bases) and byte-coded alphabets
engine = importlib.import_module(PKG + ".engine")
eng = engine.get_engine()
def rate(bases, offsets, k, code_bits, label):
    ub, uo, counts, _ = synth.dedup(bases, offsets)
    rs = eng.upload_reads(ub, uo, code_bits=code_bits)
    idx = eng.kmer_index(rs, k)
    pa, pb, _ = eng.candidate_pairs(rs, idx, k)
    lens = rs.length[:rs.n_reads].to(torch.int64)
    cells = int((lens[pa.long()] * lens[pb.long()]).sum().item())
    eng.overlap_scores(rs, pa, pb); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.overlap_scores(rs, pa, pb); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{label}: {pa.shape[0]} pairs, {cells:.3e} cells, DP {ms:.2f} ms = {cells/ms/1e6:.0f} GCUPS ({eng.dp_plan(rs.max_len)['mode'] if code_bits == 2 else 'byte-coded'})")
g = synth.random_genome(400_000, 3)
b, o = synth.simulate_reads(g, 4000, 4000, 0.02, seed=5)
rate(b, o, 12, 2, "l=4000 (anti-diagonal CTA-per-pair kernel)")
b, o = synth.simulate_reads(synth.phix_like_genome(), 20000, 150, 0.01, seed=6)
rate(b, o, 5, 2, "PhiX-like N=20000 l=150, 2-bit packed")
rate(b, o, 5, 8, "same reads, byte-coded (8-bit symbols in the packed wavefront kernel)")
