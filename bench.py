#!/usr/bin/env python
"""Benchmark of the overlap-detection hot path (BASELINE.json metric: overlap GCUPS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--k K]
    python bench.py --impl reference ...      # the CPU arm (oracle port on the host cores)

A step = one pass of the hot path over the workload: pack -> k-mer keys -> prefix index ->
candidate join -> overlap DP -> edge expansion (-> edge exchange when N > 1).
  value  GCUPS = sum over candidate pairs of len(a)*len(b) (cells as the reference fills them,
         aligners.py:33-34) / step time, inputs resident in HBM.
  e2e    the same through the host-buffer call (engine.overlap_edges): H2D of the reads and
         D2H of the edge rows inside the timed region.
Prints ONE JSON line on rank 0.  At N = 1 the line also carries `configs_extra`: the other
BASELINE.json configs (PhiX N=1000 / N=50,000, the l=1000 long-read set at k=8 and k=5, and the
experiments.py parameter sweep as one batched job), each measured the same way in the same run.
"""
from __future__ import annotations

import argparse
import contextlib
import importlib
import io
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "genome-assembly-using-overlap-graphs_b200"

METRIC = "overlap_gcups"
UNIT = "GCUPS"
OPS_PER_CELL = 7        # compare, select, 3 adds, 2 max
SCORING = "match 10, mismatch -1, indel -2^31 (reference defaults)"
L2_NOTE = "GPU arm: flushed between timed iterations (256 MiB memset)"


class Workload:
    """Host side of one workload: the unique reads (the host part of the builder, read_copies)."""

    def __init__(self, name, seed, k):
        synth = importlib.import_module(PKG + ".synth")
        bases, offsets = synth.make_workload(name, seed)
        self.name, self.seed, self.k = name, seed, k
        self.n_reads = len(offsets) - 1
        self.ub, self.uo, self.counts, _ = synth.dedup(bases, offsets)
        self.U = len(self.counts)
        self.total_bases = int(self.uo[-1])
        self.max_len = int((self.uo[1:] - self.uo[:-1]).max())
        self.has_dups = bool(self.counts.max() > 1)

    def shared_config(self, pairs=None):
        """The keys both arms print (the driver compares them)."""
        return {"workload": self.name, "k": self.k, "seed": self.seed, "reads": self.n_reads,
                "unique_reads": self.U, "max_read_len": self.max_len, "candidate_pairs": pairs, "scoring": SCORING,
                "l2": L2_NOTE}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML from a
    thread every 5 ms (nvidia-smi -lms cannot sample a region of a few tens of ms)."""

    def __init__(self, gpu_index):
        self.rows = []          # (t, sm_mhz, max_mhz, power_w, reasons bitmask)
        self.gpu_index = gpu_index
        self.stop_flag = threading.Event()
        self.thread = None
        self.err = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip() != ""]
            try:
                return int(ids[self.gpu_index])
            except Exception:
                return self.gpu_index
        return self.gpu_index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as exc:          # noqa: BLE001
            self.err = f"NVML unavailable: {exc}"
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = None
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.time(), sm, self.max_mhz, pw, rs))
            except Exception as exc:      # noqa: BLE001
                self.err = str(exc)
                return
            time.sleep(0.005)

    def stop(self, t_begin=None, t_end=None):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"]}
        rows = [r for r in self.rows if (t_begin is None or r[0] >= t_begin) and (t_end is None or r[0] <= t_end)]
        if len(rows) < 3:
            rows = self.rows
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(n for n, bit in names.items() if any(r[4] & bit for r in rows))
        power = [r[3] for r in rows if r[3] is not None]
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": rows[0][2],
                "power_w_max": max(power) if power else None, "samples": len(rows), "reasons": reasons}


# --------------------------------------------------------------------------------------- CPU arm
def cpu_arm(ub, uo, pair_a, pair_b, target_s=12.0, nthreads=0, total_pairs=None):
    """The oracle port of aligners.overlap_alignment (full matrices + traceback walk, like the
    reference) over a bounded sample of the workload's candidate pairs, on all host threads."""
    from oracle import overlap_oracle as orc
    cores = (os.cpu_count() or 1) if nthreads <= 0 else nthreads     # torchrun exports OMP_NUM_THREADS=1: be explicit
    lens = (uo[1:] - uo[:-1]).astype(np.int64)
    rng = np.random.Generator(np.random.PCG64(99))
    P = len(pair_a)
    if P == 0:
        return {"value": 0.0, "unit": UNIT, "cores": cores, "kind": "port", "sample": "no candidate pairs",
                "seconds": 0.0, "pairs": 0, "cells": 0}
    cells_per_pair = float((lens[pair_a[:4096]] * lens[pair_b[:4096]]).mean())
    probe = min(P, max(2 * cores, int(256 * cores * 22500 / max(cells_per_pair, 1.0))))
    sel = rng.choice(P, size=probe, replace=False) if P > probe else np.arange(P)
    t0 = time.perf_counter()
    orc.overlap_pairs(ub, uo, pair_a[sel], pair_b[sel], full=True, nthreads=cores)
    dt = max(time.perf_counter() - t0, 1e-6)
    n = int(min(P, max(probe, probe * target_s / dt)))
    sel = rng.choice(P, size=n, replace=False) if P > n else np.arange(P)
    cells = int((lens[pair_a[sel]] * lens[pair_b[sel]]).sum())
    t0 = time.perf_counter()
    orc.overlap_pairs(ub, uo, pair_a[sel], pair_b[sel], full=True, nthreads=cores)
    dt = time.perf_counter() - t0
    return {"value": cells / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} of {total_pairs or P} candidate pairs ({cells:.3e} cells) in {dt:.2f} s, oracle/overlap_oracle.c "
                      f"ovo_overlap_pairs(full=1), OpenMP x{cores}",
            "seconds": dt, "pairs": n, "cells": cells}


def host_candidate_pairs(ub, uo, k, max_pairs=None):
    """Candidate list on the host for the CPU arm (NumPy k-mer join, same rule as
    overlapGraphs.py:30-52).  Not timed.  With max_pairs, only a random subset of the source
    reads is expanded (all their candidates), so memory stays bounded on the big workloads.
    Returns (pair_a, pair_b, total_pairs)."""
    lens = (uo[1:] - uo[:-1]).astype(np.int64)
    code = np.zeros(256, np.uint64)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = i
    valid = np.nonzero(lens >= k)[0]
    pk = np.zeros(len(valid), np.uint64)
    sk = np.zeros(len(valid), np.uint64)
    for i in range(k):
        pk = pk * np.uint64(4) + code[ub[uo[valid] + i]]
        sk = sk * np.uint64(4) + code[ub[uo[valid + 1] - k + i]]
    order = np.argsort(pk, kind="stable")
    spk = pk[order]
    lo = np.searchsorted(spk, sk, "left")
    hi = np.searchsorted(spk, sk, "right")
    cnt = hi - lo
    total = int(cnt.sum() - (pk == sk).sum())            # a read sits in its own bucket iff prefix == suffix key
    src = np.arange(len(valid))
    if max_pairs is not None and int(cnt.sum()) > max_pairs:
        rng = np.random.Generator(np.random.PCG64(7))
        perm = rng.permutation(len(valid))
        take = np.searchsorted(np.cumsum(cnt[perm]), max_pairs) + 1
        src = np.sort(perm[:take])
    c = cnt[src]
    a = np.repeat(valid[src], c)
    starts = np.repeat(lo[src], c)
    within = np.arange(int(c.sum())) - np.repeat(np.cumsum(c) - c, c)
    b = valid[order][starts + within]
    keep = a != b
    return a[keep].astype(np.int32), b[keep].astype(np.int32), total


def numba_reference_sample(wl, pair_a, pair_b, seconds=8.0):
    """The LIVE, unmodified Numba reference (aligners.overlap_alignment) on a small sample of the same
    pairs -- only where a reference checkout is importable (the build container; the GPU box has none)."""
    try:
        from oracle import ref_loader
        if not ref_loader.available():
            return None
        ref_al, _ = ref_loader.load()
    except Exception:                      # noqa: BLE001  (numba missing, ...)
        return None
    buf = wl.ub.tobytes().decode("ascii")
    off = wl.uo.tolist()
    ref_al.overlap_alignment("ACGT", "ACGT")                # JIT warm-up (not timed)
    n, cells = 0, 0
    t0 = time.perf_counter()
    for a, b in zip(pair_a.tolist(), pair_b.tolist()):
        s, t = buf[off[a]:off[a + 1]], buf[off[b]:off[b + 1]]
        ref_al.overlap_alignment(s, t)
        n += 1
        cells += len(s) * len(t)
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return {"value": cells / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "reference (live Numba, build container only)",
            "sample": f"{n} candidate pairs ({cells:.3e} cells) in {dt:.2f} s, aligners.overlap_alignment, 1 process"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args.workload, args.seed, args.k)
    pa, pb, total_pairs = host_candidate_pairs(wl.ub, wl.uo, args.k, max_pairs=8_000_000)
    per_step = max(2.0, min(20.0, 60.0 / max(args.steps + args.warmup, 1)))
    vals, secs = [], []
    last = None
    for it in range(args.warmup + args.steps):
        last = cpu_arm(wl.ub, wl.uo, pa, pb, target_s=per_step, total_pairs=total_pairs)
        if it >= args.warmup:
            vals.append(last["value"]); secs.append(last["seconds"])
    v = float(np.mean(vals))
    lens = (wl.uo[1:] - wl.uo[:-1]).astype(np.int64)
    config = wl.shared_config(int(total_pairs))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "reference_arm_note": (f"each step = the CPU port over a random sample of the candidate pairs "
                                   f"(~{per_step:.0f} s of work); ms_per_step is that sample's duration, not a workload step")}
    line["cpu_baseline"]["value"] = v
    numba = numba_reference_sample(wl, pa[:4000], pb[:4000])
    if numba is not None:
        line["cpu_baseline_numba"] = numba
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- GPU arm
class Env:
    """Process-wide bench state: torch, the engine, ranks, the L2 flush buffer, the integer probe."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.engine_mod = importlib.import_module(PKG + ".engine")
        self.par = importlib.import_module(PKG + ".parallel")
        self.nat = importlib.import_module(PKG + "._native")
        self.eng = self.engine_mod.get_engine(self.local_rank)
        self.dev = self.eng.device
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)       # > 126 MB L2
        self._probe = None

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def probe(self):
        """Integer-pipe issue rates measured on this GPU in this run (G lane-instructions / s)."""
        if self._probe is None:
            import ctypes
            names = {0: "iadd3", 1: "imad", 2: "vimnmx_s32", 3: "viaddmnmx_s16x2", 4: "dp_mix", 5: "prmt", 6: "lop3",
                     7: "lop3_imad_pair", 8: "vimnmx3_imad_distinct_regs", 9: "dp_form1_column"}
            out = {}
            for kind, nm in names.items():
                g, ms = ctypes.c_double(), ctypes.c_double()
                self.nat.check(self.nat.lib.ovl_int_peak_probe(self.eng._ctx, kind, 2000, ctypes.byref(g), ctypes.byref(ms)))
                out[nm] = round(g.value, 1)
            self._probe = out
        return self._probe


def measure(env, wl, steps, warmup, cpu_seconds, use_peer=True, with_cpu=True, sample_clocks=True):
    """One workload through the value leg (device-resident inputs), the e2e leg (host buffers) and the
    CPU arm; returns the JSON line as a dict (rank 0) or None."""
    torch, dist, eng, par, dev = env.torch, env.dist, env.eng, env.par, env.dev
    rank, world = env.rank, env.world
    U, k = wl.U, wl.k

    # pinned host buffers (the e2e leg copies from these every step)
    h_bases = torch.from_numpy(wl.ub[:wl.total_bases].copy()).pin_memory()
    h_off = torch.from_numpy(wl.uo.copy()).pin_memory()
    h_counts = torch.from_numpy(wl.counts.copy()).pin_memory()
    node_off_np = np.zeros(U + 1, np.int64)
    np.cumsum(wl.counts, out=node_off_np[1:])

    # device-resident inputs for the `value` leg
    d_ascii = torch.empty(wl.total_bases + 64, dtype=torch.uint8, device=dev)
    d_ascii[:wl.total_bases].copy_(h_bases)
    d_off = h_off.to(dev)
    d_copies = h_counts.to(dev) if wl.has_dups else None
    d_node_off = torch.from_numpy(node_off_np).to(dev) if wl.has_dups else None

    shard = (rank, world)
    peer = None          # set after the first (NCCL-gather) pass, when the total edge count is known
    exchange = "none (1 GPU)"

    def device_step(record=None):
        # K0-K3 as one library call + one host sync (the pair / edge totals), then the pair fill
        cand = eng.build_candidates(d_ascii, d_off, U, wl.max_len, k, d_copies, d_node_off, shard)
        pa, pb = eng.fill_pairs(cand)
        if record is not None:
            record["k1"].record()           # end of the k-mer stages (K0-K3)
        # K6 is fused into the DP epilogue; with duplicate reads the row offsets come from the join index
        edges = eng.candidate_edges(cand, pa, pb,
                                    events=(record["dp0"], record["dp1"]) if record is not None else None,
                                    sink=peer.slot if peer is not None else None)
        if peer is not None:
            # the DP epilogue has stored this rank's rows straight into rank 0's buffer over NVLink
            peer.barrier()
            return cand.rs, pa, pb, None, peer.result()
        edges_all = par.gather_edges(edges, 0) if world > 1 else edges
        return cand.rs, pa, pb, edges, edges_all

    def list_hash(edges_all):
        """Order-sensitive fingerprint of the complete list (rank 0), as an unsigned 64-bit int."""
        if rank != 0 or edges_all is None:
            return 0
        return int(eng.edge_hash(edges_all).item()) & 0xFFFFFFFFFFFFFFFF

    # ---- untimed: workload statistics (cells as the reference fills them)
    rs, pa, pb, edges, edges_all = device_step()
    torch.cuda.synchronize()
    lens_d = rs.length[:U].to(torch.int64)
    cells_local = int((lens_d[pa.long()] * lens_d[pb.long()]).sum().item()) if pa.numel() else 0
    pairs_local = int(pa.shape[0])
    edges_local = int(edges.shape[0])
    stat = torch.tensor([cells_local, pairs_local, edges_local], dtype=torch.int64, device=dev)
    stat_max = torch.tensor([pairs_local, cells_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(stat)
        dist.all_reduce(stat_max, op=dist.ReduceOp.MAX)
    cells, pairs, n_edges = (int(x) for x in stat.cpu().tolist())
    pairs_rank_max, cells_rank_max = (int(x) for x in stat_max.cpu().tolist())
    edge_hash = list_hash(edges_all)
    plan = eng.dp_plan(wl.max_len)
    del rs, pa, pb, edges, edges_all, lens_d
    if world > 1:
        exchange = "NCCL send/recv gather of the per-rank edge slices"
        if use_peer:
            err = None
            try:
                peer = par.PeerEdgeBuffer(n_edges, dev)
            except Exception as exc:                      # noqa: BLE001
                err, peer = exc, None
            agree = torch.tensor([1 if peer is not None else 0], device=dev)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)  # every rank takes the same path
            if int(agree.item()) != 1:
                peer = None
                if rank == 0:
                    print(f"[bench] peer-memory path unavailable ({err}); using the NCCL gather", file=sys.stderr)
            else:
                # self-check: the peer-store path must reproduce the gathered list row for row
                _, _, _, _, chk_all = device_step()
                ok = torch.tensor([1 if (rank != 0 or list_hash(chk_all) == edge_hash) else 0], device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()) != 1:
                    raise SystemExit("peer-store edge list differs from the gathered one")
                exchange = "DP epilogue stores edge rows directly into rank 0's HBM (NVLink peer memory)"

    # ---- value leg: inputs resident in HBM
    sampler = ClockSampler(env.local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
    for _ in range(warmup):
        device_step()
        env.flush.zero_()
    env.barrier()
    t_begin = time.time()
    step_ms, dp_ms, kmer_ms, off_ms = [], [], [], []
    launches0 = eng.launches
    for _ in range(steps):
        env.flush.zero_()                               # flush L2 between timed iterations
        rec = {"dp0": env.ev(), "dp1": env.ev(), "k1": env.ev()}
        e0, e1 = env.ev(), env.ev()
        env.barrier()
        e0.record()
        device_step(rec)
        e1.record()
        env.barrier()
        step_ms.append(e0.elapsed_time(e1))
        dp_ms.append(rec["dp0"].elapsed_time(rec["dp1"]))
        kmer_ms.append(e0.elapsed_time(rec["k1"]))
        off_ms.append(rec["k1"].elapsed_time(rec["dp0"]))
    launches = eng.launches - launches0
    clocks = sampler.stop(t_begin, time.time()) if sampler else None
    t = torch.tensor([sum(step_ms), sum(dp_ms), sum(kmer_ms), sum(off_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)        # max over ranks
    total_ms, dp_total_ms, kmer_total_ms, off_total_ms = (float(x) for x in t.cpu().tolist())
    ms_per_step = total_ms / steps
    value = cells / (ms_per_step * 1e-3) / 1e9

    # ---- e2e leg: host buffers in, host edge rows out, every step
    sink = par.SharedEdgeSink(initial_rows=n_edges) if world > 1 else None

    def e2e_step():
        if world == 1:
            return eng.overlap_edges(h_bases, h_off, h_counts if wl.has_dups else None, k, shard, reuse_host_buffer=True)
        # every rank copies its slice over its own PCIe link into one shared, page-locked host buffer;
        # after the barrier rank 0 holds the complete ordered edge list in host memory
        # the inputs cross PCIe once in total: rank r uploads its 1/world, one all-gather over NVLink completes them
        eng.overlap_edges(h_bases, h_off, h_counts if wl.has_dups else None, k, shard, host_sink=sink,
                          upload_group=dist.group.WORLD)
        dist.barrier()
        return sink.rows()

    for _ in range(max(1, min(warmup, 2))):           # at least one: the first call allocates the pinned result buffer
        e2e_step()
    e2e_ms = []
    d2h_bytes = 0
    out = None
    for _ in range(steps):
        env.flush.zero_()
        env.barrier()
        t0 = time.perf_counter()
        out = e2e_step()
        torch.cuda.synchronize()
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
        d2h_bytes = out.nbytes + 16
    # untimed: the host result of the last e2e step must be, row for row, the device-resident step's edge list
    e2e_ok = bool(out.shape[0] == n_edges and eng.edge_hash_host(out) == edge_hash) if rank == 0 else None
    te = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_per_step = float(te.item()) / steps
    h2d_bytes = wl.total_bases + h_off.numel() * 8 + (h_counts.numel() * 4 + (U + 1) * 8 if wl.has_dups else 0)

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (the DP): integer pipes, not HBM
        probe = env.probe()
        dp_ms_avg = dp_total_ms / steps
        cells_per_launch = cells_rank_max                      # the slowest rank's slice (max-over-ranks time)
        achieved = cells_per_launch * OPS_PER_CELL / (dp_ms_avg * 1e-3) / 1e12
        # Peak of the integer pipes for this op class, measured in this run: the ALU pipe issues
        # VIADDMNMX.U16x2 at `viaddmnmx_s16x2` G lane-instr/s, each doing 4 algorithmic 16-bit ops
        # (2 halves x (add + min)); the FMA pipe co-issues IMAD at `imad` G lane-instr/s, each a
        # packed add = 2 algorithmic ops.
        packed = plan["mode"] == "packed16"
        peak = (probe["viaddmnmx_s16x2"] * 4 + probe["imad"] * 2) / 1e3 if packed \
            else (probe["viaddmnmx_s16x2"] * 2 + probe["imad"]) / 1e3
        # SURVEY 8(d)'s own denominator: plain int32 lanes x clock, one algorithmic op per lane-instruction, at the
        # rate the min/max instruction issues (VIMNMX.S32: 128 lanes/clk/SM).  Packed DPX does two 16-bit cells per
        # lane-instruction, so this fraction exceeds 1 by design; it is reported, not used as the ceiling.
        peak_int32 = probe["vimnmx_s32"] / 1e3
        pk, pk_kind = peaks()
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "dp_traffic.json")) as fh:
                ent = json.load(fh).get(f"{wl.name}:k{k}:gpus{world}")
            if ent:
                traffic = ent["dram_bytes_read"] + ent["dram_bytes_write"]      # one ncu capture, per launch
        except Exception:                                  # noqa: BLE001
            traffic = None
        kernel = {"packed16": f"overlap_dp_kernel<{plan['lanes']},{plan['cols']},packed16>",
                  "int32": f"overlap_dp_kernel<{plan['lanes']},{plan['cols']},int32>",
                  "long-read": "overlap_dp_long_kernel"}[plan["mode"]]
        roofline = {"bound": "int-pipe (ALU DPX + FMA IMAD)", "kernel": kernel,
                    "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak if peak else None,
                    "traffic": traffic, "ops_per_cell": OPS_PER_CELL,
                    "dp_gcups": cells_per_launch / (dp_ms_avg * 1e-3) / 1e9, "dp_ms": dp_ms_avg,
                    "peak_source": "ovl_int_peak_probe in this run: 4 ops x VIADDMNMX.16x2 rate + 2 ops x IMAD rate",
                    "frac_vs_int32_lanes": achieved / peak_int32, "peak_int32_lanes": peak_int32,
                    "peak_int32_lanes_source": "VIMNMX.S32 issue rate measured in this run (SURVEY 8d's denominator)",
                    # what a pure stream of the kernel's own form-1 columns (PRMT, IMAD, 2x VIADDMNMX on distinct
                    # registers, no loop overhead) sustains on this GPU: 2 cells per 4 lane-instructions
                    "same_mix_stream_gcups": probe["dp_form1_column"] / 4 * 2,
                    "int_probe_gops": probe, "hbm_peak_gbs": pk.get("hbm_gbs"), "hbm_peak_source": pk_kind}
        # ---- the k-mer stages (K0-K3) and edge expansion (K6): HBM-bound; algorithmic bytes per SURVEY 8(d).
        # Every rank packs / indexes / counts all reads (replicated) and fills ITS slice of the pair list.
        # K0 pack: ASCII in + packed out; K1 (fused into K0): 16 B of keys out; K2: 12 B x (1 + 2 x passes) with
        # 10-bit digits (one pass for k <= 5); K3: 16 B per read + 12 B per pair
        passes = (2 * k + 9) // 10
        kb = (wl.total_bases + wl.total_bases / 4) + 16 * U + 12 * (1 + 2 * passes) * U + 16 * U + 12 * pairs_rank_max
        k_ms = kmer_total_ms / steps
        o_ms = off_total_ms / steps
        hbm = pk.get("hbm_gbs")
        k6_bytes = 16 * pairs_rank_max + 16 * (n_edges / world)
        kmer = {"algorithmic_bytes_k0_k3": int(kb), "ms": k_ms,
                "k0_k3_gbs": kb / (k_ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm,
                "k0_k3_frac": kb / (k_ms * 1e-3) / 1e9 / hbm,
                "k6": {"edge_offset_stage_ms": o_ms, "algorithmic_bytes": int(k6_bytes),
                       "note": "the edge rows are written by the DP epilogue (their time is inside dp_ms); "
                               "edge_offset_stage_ms is what precedes the DP when reads have copies"},
                "k0_k3_plus_k6_offsets_frac": kb / ((k_ms + o_ms) * 1e-3) / 1e9 / hbm,
                "per_rank": world > 1,
                "note": "pack+keys, index (+ bucket table), join count / scan / totals (one library call), the host sync on "
                        "the totals, pair fill; timed with CUDA events inside the step"}
        # ---- CPU baseline on this box's host cores (bounded sample)
        cpu = numba = None
        if with_cpu and world == 1 and pairs > 0:
            # a bounded random sample of the SAME candidate list the GPU just processed
            rs_, pa_, pb_, _, _ = device_step()
            n_s = min(pairs, 4_000_000)
            idx = torch.randperm(pairs, device=dev)[:n_s] if pairs > n_s else torch.arange(pairs, device=dev)
            pa_h, pb_h = pa_[idx].cpu().numpy(), pb_[idx].cpu().numpy()
            del rs_, pa_, pb_, idx
            cpu = cpu_arm(wl.ub, wl.uo, pa_h, pb_h, target_s=cpu_seconds, total_pairs=pairs)
            cpu = {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
            numba = numba_reference_sample(wl, pa_h[:4000], pb_h[:4000], seconds=min(cpu_seconds, 8.0))
        config = wl.shared_config(pairs)            # identical keys and values in the reference arm's line
        stats = {"edges": n_edges, "cells": cells, "edge_list_hash": f"{edge_hash:016x}",
                 "sharding": f"pair-range x{world}", "exchange": exchange}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "int16x2" if packed else "int32", "data": "synthetic",
                "config": config, "workload_stats": stats,
                "pairs_per_s": pairs_rank_max / (k_ms * 1e-3) * world,
                "stage_ms": {"kmer_index_join": k_ms, "edge_offsets": o_ms, "overlap_dp_fused_expand": dp_ms_avg,
                             "exchange_and_rest": ms_per_step - k_ms - o_ms - dp_ms_avg},
                "kmer_stages": kmer,
                "e2e": {"value": cells / (e2e_per_step * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_per_step,
                        "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                        "edge_list_hash_matches_device_path": e2e_ok,
                        "path": "engine.overlap_edges: pinned host reads in, host edge rows out, D2H overlapped with the DP"
                                + ("; the reads cross PCIe once in total (each rank uploads 1/N, one all-gather over NVLink); each rank "
                                   "writes its slice into one shared page-locked host buffer" if world > 1 else "")},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        if numba is not None:
            line["cpu_baseline_numba"] = numba
    if world > 1:
        sink.close()
        dist.barrier()
    return line


# --------------------------------------------------------------------------------------- the other configs
def _brief(line):
    """The part of a measured line that configs_extra keeps."""
    r = line["roofline"]
    return {"value": line["value"], "unit": UNIT, "ms_per_step": line["ms_per_step"], "steps": line["steps"],
            "warmup": line["warmup"], "dtype": line["dtype"],
            "config": {**{kk: line["config"][kk] for kk in ("workload", "k", "reads", "unique_reads", "max_read_len",
                                                            "candidate_pairs")},
                       **{kk: line["workload_stats"][kk] for kk in ("edges", "cells", "edge_list_hash")}},
            "e2e": {kk: line["e2e"][kk] for kk in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step",
                                                   "edge_list_hash_matches_device_path")},
            "stage_ms": line["stage_ms"],
            "k0_k3_frac": line["kmer_stages"]["k0_k3_frac"],
            "roofline": {kk: r[kk] for kk in ("kernel", "bound", "achieved", "peak", "unit", "frac", "dp_gcups", "dp_ms")},
            "cpu_baseline": line["cpu_baseline"], "gpu_launches": line["gpu_launches"]}


def dropin_wall(env, wl, reps=3):
    """The true drop-in call: list[str] -> (nx.DiGraph, read_copies), wall clock (host dedup, H2D, kernels,
    D2H, NetworkX construction)."""
    synth = importlib.import_module(PKG + ".synth")
    og = importlib.import_module(PKG + ".overlapGraphs")
    bases, offsets = synth.make_workload(wl.name, wl.seed)
    reads = synth.to_strings(bases, offsets)
    og.construct_overlap_graph_nx_k(reads, k=wl.k)                    # warm-up
    ts = []
    G = None
    for _ in range(reps):
        env.torch.cuda.synchronize()
        t0 = time.perf_counter()
        G, rc = og.construct_overlap_graph_nx_k(reads, k=wl.k)
        ts.append(time.perf_counter() - t0)
    return {"call": "construct_overlap_graph_nx_k(list[str], k) -> (nx.DiGraph, read_copies)",
            "wall_ms": min(ts) * 1e3, "wall_ms_median": statistics.median(ts) * 1e3, "reps": reps,
            "nodes": G.number_of_nodes(), "edges": G.number_of_edges()}


def sweep_sets(iterations, seed0=1000, eng=None):
    """The experiments.py:49-53 grid (N x l x p), `iterations` read sets per point, on the PhiX-like genome.
    With an engine the reads come from the DEVICE simulator (ovl_simulate_reads; same bytes as its NumPy mirror)."""
    synth = importlib.import_module(PKG + ".synth")
    g = synth.phix_like_genome()
    g_dev = eng._to_device(g, eng.torch_uint8()) if eng is not None else None
    sets, meta = [], []
    i = 0
    for it in range(iterations):
        for n in (100, 316, 1000, 3162, 10000):
            for l in (50, 100, 150):
                for p in (0.001, 0.01, 0.1):
                    if eng is not None:
                        a_dev, o_dev = eng.simulate_reads(g_dev, n, l, p, seed0 + i)
                        o = o_dev.cpu().numpy()
                        b = a_dev[:int(o[-1])].cpu().numpy()
                    else:
                        b, o = synth.simulate_reads_counter(g, n, l, p, seed0 + i)
                    i += 1
                    sets.append(synth.to_strings(b, o))
                    meta.append((n, l, p, it))
    return sets, meta


def _sweep_cpu_worker(job):
    from oracle import overlap_oracle as orc
    reads, k = job
    if reads is None:                       # pool warm-up: load the oracle library in this worker
        orc.overlap_alignment("ACGT", "CGTA")
        return 0
    nodes, edges, rc = orc.construct_overlap_graph(reads, k, nthreads=1, full=True)
    return len(edges)


def sweep_extra(env, iterations=10, cpu_seconds=20.0):
    """BASELINE.json configs[4]: every graph build of the reference's parameter sweep (experiments.py:49-53:
    N x l x p x k x 10 iterations) as ONE batched GPU job per k, against the CPU port run the reference's way --
    one process per core over the parameter sets (experiments.py:537)."""
    import multiprocessing as mp
    og = importlib.import_module(PKG + ".overlapGraphs")
    torch = env.torch
    t0 = time.perf_counter()
    sets, meta = sweep_sets(iterations, eng=env.eng)
    t_sim = time.perf_counter() - t0
    n_reads = sum(len(s) for s in sets)
    out = {"grid": "N in {100,316,1000,3162,10000} x l in {50,100,150} x p in {0.001,0.01,0.1} x k in {5,10,15} "
                   f"x {iterations} iterations (experiments.py:49-53), PhiX-like genome",
           "read_sets_per_k": len(sets), "reads_per_k": n_reads,
           "read_simulation_s": t_sim, "read_simulation": "device (ovl_simulate_reads) + D2H + Python strings, untimed part of the job",
           "per_k": {}}
    one_iter = [s for s, m in zip(sets, meta) if m[3] == 0]
    cores = os.cpu_count() or 1
    total_gpu_s = total_graph_s = 0.0
    cpu_s_one_iter = 0.0
    # worker processes are spawned (CUDA and OpenMP state must not be forked) and warmed before any clock starts
    pool = mp.get_context("spawn").Pool(cores)
    pool.map(_sweep_cpu_worker, [(None, 0)] * (4 * cores), chunksize=1)
    for k in (5, 10, 15):
        og.overlap_edge_rows_batch(sets[:8], k)                        # warm-up
        launches0 = env.eng.launches
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows = og.overlap_edge_rows_batch(sets, k)                     # list[str] sets in, host edge rows per set out
        torch.cuda.synchronize()
        t_rows = time.perf_counter() - t0
        launches = env.eng.launches - launches0
        n_edges = int(sum(r[3].shape[0] for r in rows))
        del rows
        t0 = time.perf_counter()
        graphs = og.construct_overlap_graphs_batch(one_iter, k)        # the full drop-in result for one iteration
        t_graph = time.perf_counter() - t0
        g_edges = int(sum(g.number_of_edges() for g, _ in graphs))
        del graphs
        # CPU arm: the oracle builder, one process per core over the sets of ONE iteration (bounded sample)
        jobs = sorted(((s, k) for s in one_iter), key=lambda j: -len(j[0]))     # largest first: balanced tail
        t0 = time.perf_counter()
        cpu_edges = sum(pool.imap_unordered(_sweep_cpu_worker, jobs, chunksize=1))
        t_cpu = time.perf_counter() - t0
        assert cpu_edges == g_edges, (cpu_edges, g_edges)
        out["per_k"][str(k)] = {"gpu_edge_rows_all_sets_s": t_rows, "edges_all_sets": n_edges, "gpu_launches": int(launches),
                                "gpu_graphs_one_iteration_s": t_graph, "edges_one_iteration": g_edges,
                                "cpu_port_one_iteration_s": t_cpu}
        total_gpu_s += t_rows
        total_graph_s += t_graph
        cpu_s_one_iter += t_cpu
    pool.close()
    pool.join()
    out["gpu_job_s"] = total_gpu_s
    out["gpu_job_note"] = ("construct-ready host edge rows for all read sets of the sweep (3 batched jobs, one per k): host "
                          "dedup + H2D + kernels + D2H, wall clock")
    out["gpu_graphs_one_iteration_s"] = total_graph_s
    out["cpu_baseline"] = {"value": cpu_s_one_iter * iterations, "unit": "s (extrapolated: measured one iteration x "
                           f"{iterations})", "measured_one_iteration_s": cpu_s_one_iter, "cores": cores, "kind": "port",
                           "sample": "oracle construct_overlap_graph (C DP, full matrices + traceback) over the "
                                     f"{len(one_iter)} read sets of one iteration per k, one process per core (spawned pool, warm), "
                                     "edge count checked against the GPU graphs"}
    return out


def next_rows_extra(env):
    """SURVEY 8(f) rows f1, f2, f4 measured through their drop-in calls, each result compared with the oracle
    (the checker) in the same run: contig -> genome local alignment (aligners.py:85-202 as the reference's
    performanceMeasures.py:219-221 loop uses it), the all-pairs builder (overlapGraphs.py:196-232) and the
    cycle removal (overlapGraphs.py:106-130)."""
    import numpy as np
    from oracle import overlap_oracle as orc
    synth = importlib.import_module(PKG + ".synth")
    al = importlib.import_module(PKG + ".aligners")
    og = importlib.import_module(PKG + ".overlapGraphs")
    torch = env.torch
    out = {}
    genome_u8 = synth.phix_like_genome()
    genome = genome_u8.tobytes().decode()
    G = len(genome)
    rng = np.random.Generator(np.random.PCG64(2024))
    # f1: 96 contigs (100 .. 1,000 bases, 1 % substitutions) against the 5,386-base genome
    contigs = []
    for i in range(96):
        L = int(rng.integers(100, 1001))
        st = int(rng.integers(0, G - L))
        c = genome_u8[st:st + L].copy()
        flip = rng.random(L) < 0.01
        c[flip] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(flip.sum()))]
        contigs.append(c.tobytes().decode())
    cells = sum(len(c) for c in contigs) * G
    al.local_alignment_batch(contigs[:4], genome)                      # warm-up
    t_batch = 1e30
    for _ in range(3):                                                 # best of three: the first full-size call also sizes the arenas
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = al.local_alignment_batch(contigs, genome)
        t_batch = min(t_batch, time.perf_counter() - t0)
    t0 = time.perf_counter()
    one = [al.local_alignment(c, genome) for c in contigs[:16]]
    t_one = (time.perf_counter() - t0) / 16
    t0 = time.perf_counter()
    ref = [orc.local_alignment(c, genome) for c in contigs[:16]]
    t_cpu = (time.perf_counter() - t0) / 16
    assert got[:16] == ref and one == ref
    out["f1_local_alignment"] = {
        "contigs": len(contigs), "genome": G, "cells": cells,
        "batch_one_launch_ms": t_batch * 1e3, "batch_ms_per_contig": t_batch * 1e3 / len(contigs),
        "batch_mcups": cells / t_batch / 1e6, "per_call_ms_per_contig": t_one * 1e3,
        "oracle_c_port_ms_per_contig_1_thread": t_cpu * 1e3, "reference_numba_ms_per_contig": 304.0,
        "note": "str contigs in, the reference's 6-tuples out (traceback strings included), wall clock; first 16 results "
                "compared with the oracle; 304 ms per contig is BASELINE.md's figure for the Numba reference"}
    # f2: the all-pairs builder on 400 reads (159,600 ordered pairs)
    b, o = synth.simulate_reads(genome_u8, 400, 100, 0.01, seed=11)
    reads = synth.to_strings(b, o)
    with contextlib.redirect_stdout(io.StringIO()):
        og.construct_overlap_graph_string(reads[:50])                  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g_gpu = og.construct_overlap_graph_string(reads)
        t_gpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        g_cpu = orc.construct_overlap_graph_string(reads, nthreads=0)
        t_cpu = time.perf_counter() - t0
    g_gpu = g_gpu[0]
    cpu_edges = g_cpu[1]                                               # (nodes, edges, read_copies)
    assert len(cpu_edges) == g_gpu.number_of_edges()
    out["f2_all_pairs_builder"] = {"reads": len(reads), "ordered_pairs_aligned": len(set(reads)) * (len(set(reads)) - 1),
                                   "edges_score_gt_0": g_gpu.number_of_edges(), "gpu_dropin_wall_ms": t_gpu * 1e3,
                                   "oracle_c_port_wall_ms_all_threads": t_cpu * 1e3, "same_edge_count_as_oracle": True}
    # f4: cycle removal on the N = 1,000 graph (same removals as the literal NetworkX restatement)
    b, o = synth.simulate_reads(genome_u8, 1000, 100, 0.01, seed=12)
    reads = synth.to_strings(b, o)
    g1, _ = og.construct_overlap_graph_nx_k(reads, k=5)
    g2 = g1.copy()
    e_before = g1.number_of_edges()
    t0 = time.perf_counter()
    og.remove_cycles_from_graph(g1)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc.remove_cycles_from_graph(g2)
    t_cpu = time.perf_counter() - t0
    assert list(g1.edges(data=True)) == list(g2.edges(data=True))
    out["f4_cycle_removal"] = {"edges_before": e_before, "edges_removed": e_before - g1.number_of_edges(),
                               "dropin_with_device_sink_peeling_ms": t_gpu * 1e3,
                               "literal_networkx_restatement_ms": t_cpu * 1e3, "same_graph_after": True}
    return out


def configs_extra(env, args):
    extra = {}
    t_all = time.perf_counter()
    plan = [("configs[0]", "phix_n1000_l100", 5, 5, 3, 3.0, True),
            ("configs[1]", "phix_n50000_l150", 5, 5, 3, 5.0, True),
            ("configs[3]_k8", "ecoli_n200k_l1000", 8, 3, 3, 5.0, False),
            ("configs[3]_k5", "ecoli_n200k_l1000", 5, 2, 3, 6.0, False)]
    for key, name, k, steps, warmup, cpu_s, dropin in plan:
        t0 = time.perf_counter()
        try:
            wl = Workload(name, args.seed, k)
            line = measure(env, wl, steps, warmup, cpu_s, sample_clocks=False)
            ent = _brief(line)
            if dropin:
                ent["dropin"] = dropin_wall(env, wl)
                if key == "configs[0]":
                    ent["dropin"]["reference_numba_s"] = 1.29      # BASELINE.md section 2 (survey container, 1 core)
            if "cpu_baseline_numba" in line:
                ent["cpu_baseline_numba"] = line["cpu_baseline_numba"]
            ent["bench_seconds"] = time.perf_counter() - t0
            extra[key] = ent
            del wl
        except Exception as exc:                           # noqa: BLE001  -- an extra must never lose the headline line
            extra[key] = {"error": f"{type(exc).__name__}: {exc}"}
        env.torch.cuda.empty_cache()
    try:
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            extra["configs[4]_sweep"] = sweep_extra(env, iterations=args.sweep_iterations)
        extra["configs[4]_sweep"]["bench_seconds"] = time.perf_counter() - t0
    except Exception as exc:                               # noqa: BLE001
        extra["configs[4]_sweep"] = {"error": f"{type(exc).__name__}: {exc}"}
    try:
        t0 = time.perf_counter()
        extra["next_rows"] = next_rows_extra(env)
        extra["next_rows"]["bench_seconds"] = time.perf_counter() - t0
    except Exception as exc:                               # noqa: BLE001
        extra["next_rows"] = {"error": f"{type(exc).__name__}: {exc}"}
    extra["bench_seconds"] = time.perf_counter() - t_all
    return extra


def run_ours(args):
    env = Env()
    wl = Workload(args.workload, args.seed, args.k)
    line = measure(env, wl, args.steps, args.warmup, args.cpu_seconds, use_peer=not args.no_peer_stores,
                   with_cpu=not args.no_cpu_baseline)
    del wl
    if env.rank == 0 and env.world == 1 and not args.no_extras:
        env.torch.cuda.empty_cache()
        line["configs_extra"] = configs_extra(env, args)
    if env.rank == 0:
        print(json.dumps(line), flush=True)
    if env.world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # BASELINE.json quotes the metric "at 1/2/4/8 B200" on configs[2] (4.6 Mb genome, 1M reads, l=150,
    # p=0.005, sharded by read-ID range); it fits one GPU (~32 GB), so it is the workload at every N.
    # The other configs are measured in the same run into `configs_extra` (N = 1).
    ap.add_argument("--workload", default="ecoli_n1m_l150")
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip configs_extra (the other BASELINE configs)")
    ap.add_argument("--sweep-iterations", type=int, default=10)
    ap.add_argument("--no-peer-stores", action="store_true", help="N>1: gather with NCCL instead of peer-memory stores")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
