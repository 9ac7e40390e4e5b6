#!/usr/bin/env python
"""Benchmark of the overlap-detection hot path (BASELINE.json metric: overlap GCUPS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--k K]
    python bench.py --impl reference ...      # the CPU arm (oracle port on the host cores)

A step = one pass of the hot path over the workload: pack -> k-mer keys -> prefix index ->
candidate join -> overlap DP -> edge expansion (-> edge gather when N > 1).
  value  GCUPS = sum over candidate pairs of len(a)*len(b) (cells as the reference fills them,
         aligners.py:33-34) / step time, inputs resident in HBM.
  e2e    the same through the host-buffer call (engine.overlap_edges): H2D of the reads and
         D2H of the edge rows inside the timed region.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
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


def load_workload(name, seed):
    synth = importlib.import_module(PKG + ".synth")
    bases, offsets = synth.make_workload(name, seed)
    ub, uo, counts, _ = synth.dedup(bases, offsets)           # host part of the builder (read_copies)
    return ub, uo, counts, len(offsets) - 1


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
    probe = min(P, 256 * cores)
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
    U = len(lens)
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ub, uo, counts, n_reads = load_workload(args.workload, args.seed)
    pa, pb, total_pairs = host_candidate_pairs(ub, uo, args.k, max_pairs=8_000_000)
    per_step = max(2.0, min(20.0, 60.0 / max(args.steps + args.warmup, 1)))
    vals, secs = [], []
    last = None
    for it in range(args.warmup + args.steps):
        last = cpu_arm(ub, uo, pa, pb, target_s=per_step, total_pairs=total_pairs)
        if it >= args.warmup:
            vals.append(last["value"]); secs.append(last["seconds"])
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": args.workload, "k": args.k, "seed": args.seed, "unique_reads": int(len(counts)),
                       "candidate_pairs": int(total_pairs)},
            "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    line["cpu_baseline"]["value"] = v
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    engine_mod = importlib.import_module(PKG + ".engine")
    par = importlib.import_module(PKG + ".parallel")
    eng = engine_mod.get_engine(local_rank)
    dev = eng.device

    ub, uo, counts, n_reads = load_workload(args.workload, args.seed)
    U = len(counts)
    total_bases = int(uo[-1])
    max_len = int((uo[1:] - uo[:-1]).max())
    has_dups = bool(counts.max() > 1)

    # pinned host buffers (the e2e leg copies from these every step)
    h_bases = torch.from_numpy(ub[:total_bases].copy()).pin_memory()
    h_off = torch.from_numpy(uo.copy()).pin_memory()
    h_counts = torch.from_numpy(counts.copy()).pin_memory()
    node_off_np = np.zeros(U + 1, np.int64)
    np.cumsum(counts, out=node_off_np[1:])

    # device-resident inputs for the `value` leg
    d_ascii = torch.empty(total_bases + 64, dtype=torch.uint8, device=dev)
    d_ascii[:total_bases].copy_(h_bases)
    d_off = h_off.to(dev)
    d_copies = h_counts.to(dev) if has_dups else None
    d_node_off = torch.from_numpy(node_off_np).to(dev) if has_dups else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    shard = (rank, world)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    peer = None          # set after the first (NCCL-gather) pass, when the total edge count is known
    exchange = "none (1 GPU)"

    def device_step(record=None):
        rs = eng.pack_reads(d_ascii, d_off, U, max_len)
        index = eng.kmer_index(rs, args.k) if args.k > 0 else None
        pa, pb, _ = eng.candidate_pairs(rs, index, args.k, shard)
        if record is not None:
            record["k1"].record()           # end of the k-mer stages (K0-K3)
        # K6 is fused into the DP epilogue; with duplicate reads a scan of the per-pair edge counts
        # (and one host read of the total) comes first
        edges = eng.overlap_edges_fused(rs, pa, pb, d_copies, d_node_off,
                                        events=(record["dp0"], record["dp1"]) if record is not None else None,
                                        sink=peer.slot if peer is not None else None)
        if peer is not None:
            # the DP epilogue has stored this rank's rows straight into rank 0's buffer over NVLink
            peer.barrier()
            return rs, pa, pb, None, peer.result()
        if world > 1:
            edges_all = par.gather_edges(edges, 0)
        else:
            edges_all = edges
        return rs, pa, pb, edges, edges_all

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- untimed: workload statistics (cells as the reference fills them)
    rs, pa, pb, edges, edges_all = device_step()
    torch.cuda.synchronize()
    eng.check_alphabet(rs)
    lens_d = rs.length[:U].to(torch.int64)
    cells_local = int((lens_d[pa.long()] * lens_d[pb.long()]).sum().item()) if pa.numel() else 0
    pairs_local = int(pa.shape[0])
    edges_local = int(edges.shape[0])
    stat = torch.tensor([cells_local, pairs_local, edges_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(stat)
    cells, pairs, n_edges = (int(x) for x in stat.cpu().tolist())
    checksum = int(edges_all.to(torch.int64).sum().item()) if (rank == 0 and edges_all is not None) else 0
    plan = eng.dp_plan(max_len)
    del rs, pa, pb, edges, edges_all
    if world > 1:
        exchange = "NCCL send/recv gather of the per-rank edge slices"
        if not args.no_peer_stores:
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
                # self-check: the peer-store path must reproduce the gathered list
                _, _, _, _, chk_all = device_step()
                ok = torch.tensor([1 if (rank != 0 or int(chk_all.to(torch.int64).sum().item()) == checksum) else 0],
                                  device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()) != 1:
                    raise SystemExit("peer-store edge list differs from the gathered one")
                exchange = "DP epilogue stores edge rows directly into rank 0's HBM (NVLink peer memory)"

    # ---- value leg: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        device_step()
        flush.zero_()
    barrier()
    t_begin = time.time()
    step_ms, dp_ms, kmer_ms = [], [], []
    launches0 = eng.launches
    for _ in range(args.steps):
        flush.zero_()                                   # flush L2 between timed iterations
        rec = {"dp0": ev(), "dp1": ev(), "k1": ev()}
        e0, e1 = ev(), ev()
        barrier()
        e0.record()
        device_step(rec)
        e1.record()
        barrier()
        step_ms.append(e0.elapsed_time(e1))
        dp_ms.append(rec["dp0"].elapsed_time(rec["dp1"]))
        kmer_ms.append(e0.elapsed_time(rec["k1"]))
    launches = eng.launches - launches0
    clocks = sampler.stop(t_begin, time.time())
    t = torch.tensor([sum(step_ms), sum(dp_ms), sum(kmer_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)        # max over ranks
    total_ms, dp_total_ms, kmer_total_ms = (float(x) for x in t.cpu().tolist())
    ms_per_step = total_ms / args.steps
    value = cells / (ms_per_step * 1e-3) / 1e9

    # ---- e2e leg: host buffers in, host edge rows out, every step
    sink = par.SharedEdgeSink(initial_rows=n_edges) if world > 1 else None

    def e2e_step():
        if world == 1:
            return eng.overlap_edges(h_bases, h_off, h_counts if has_dups else None, args.k, shard, reuse_host_buffer=True)
        # every rank copies its slice over its own PCIe link into one shared, page-locked host buffer;
        # after the barrier rank 0 holds the complete ordered edge list in host memory
        eng.overlap_edges(h_bases, h_off, h_counts if has_dups else None, args.k, shard, host_sink=sink)
        dist.barrier()
        return sink.rows()

    for _ in range(max(1, min(args.warmup, 2))):      # at least one: the first call allocates the pinned result buffer
        e2e_step()
    e2e_ms = []
    d2h_bytes = 0
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        out = e2e_step()
        torch.cuda.synchronize()
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
        d2h_bytes = out.nbytes + 16
    # untimed: the host result of the last e2e step must be the same edge list as the device-resident step's
    e2e_ok = bool(int(out.sum(dtype=np.int64)) == checksum and out.shape[0] == n_edges) if rank == 0 else None
    te = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_per_step = float(te.item()) / args.steps
    h2d_bytes = total_bases + h_off.numel() * 8 + (h_counts.numel() * 4 + (U + 1) * 8 if has_dups else 0)

    if rank == 0:
        # ---- roofline of the dominant kernel (the DP): integer pipe, not HBM
        probe = {}
        names = {0: "iadd3", 1: "imad", 2: "vimnmx_s32", 3: "viaddmnmx_s16x2", 4: "dp_mix", 5: "prmt", 6: "lop3",
                 7: "lop3_imad_pair", 8: "vimnmx3_imad_distinct_regs", 9: "dp_form1_column"}
        import ctypes
        nat = importlib.import_module(PKG + "._native")
        for kind, nm in names.items():
            g, ms = ctypes.c_double(), ctypes.c_double()
            nat.check(nat.lib.ovl_int_peak_probe(eng._ctx, kind, 2000, ctypes.byref(g), ctypes.byref(ms)))
            probe[nm] = round(g.value, 1)
        dp_ms_avg = dp_total_ms / args.steps
        cells_per_launch = cells / world                       # each rank launches the DP on its slice
        achieved = cells_per_launch * OPS_PER_CELL / (dp_ms_avg * 1e-3) / 1e12
        # Peak of the integer pipes for this op class, measured in this run: the ALU pipe issues
        # VIADDMNMX.U16x2 at `viaddmnmx_s16x2` G lane-instr/s, each doing 4 algorithmic 16-bit ops
        # (2 halves x (add + min)); the FMA pipe co-issues IMAD at `imad` G lane-instr/s, each a
        # packed add = 2 algorithmic ops.
        peak = (probe["viaddmnmx_s16x2"] * 4 + probe["imad"] * 2) / 1e3 if plan["mode"] == "packed16" \
            else (probe["viaddmnmx_s16x2"] * 2 + probe["imad"]) / 1e3
        peak_int32 = probe["lop3"] / 1e3            # int32 lanes x clock (one op per lane-instr)
        pk, pk_kind = peaks()
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "dp_traffic.json")) as fh:
                ent = json.load(fh).get(f"{args.workload}:k{args.k}:gpus{world}")
            if ent:
                traffic = ent["dram_bytes_read"] + ent["dram_bytes_write"]      # one ncu capture, per launch
        except Exception:
            traffic = None
        roofline = {"bound": "int-pipe (ALU DPX + FMA IMAD)",
                    "kernel": f"overlap_dp_kernel<{plan['lanes']},{plan['cols']},{plan['mode']}>",
                    "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak if peak else None,
                    "traffic": traffic, "ops_per_cell": OPS_PER_CELL,
                    "dp_gcups": cells_per_launch / (dp_ms_avg * 1e-3) / 1e9, "dp_ms": dp_ms_avg,
                    "peak_source": "ovl_int_peak_probe in this run: 4 ops x VIADDMNMX.16x2 rate + 2 ops x IMAD rate",
                    "frac_vs_int32_lanes": achieved / peak_int32, "peak_int32_lanes": peak_int32,
                    # what a pure stream of the kernel's own form-1 columns (PRMT, IMAD, 2x VIADDMNMX on distinct
                    # registers, no loop overhead) sustains on this GPU: 2 cells per 4 lane-instructions
                    "same_mix_stream_gcups": probe["dp_form1_column"] / 4 * 2,
                    "int_probe_gops": probe, "hbm_peak_gbs": pk.get("hbm_gbs"), "hbm_peak_source": pk_kind}
        # ---- the k-mer stages (K0-K3) and edge expansion (K6): HBM-bound; algorithmic bytes per SURVEY 8(d)
        passes = (2 * args.k + 7) // 8
        kb = (total_bases + total_bases / 4) + 48 * U + 12 * (1 + 2 * passes) * U + 16 * U + 12 * pairs
        k_ms = kmer_total_ms / args.steps
        kmer = {"algorithmic_bytes_k0_k3": int(kb), "ms": k_ms,
                "k0_k3_gbs": kb / (k_ms * 1e-3) / 1e9,
                "hbm_peak_gbs": pk.get("hbm_gbs"), "k0_k3_frac": kb / (k_ms * 1e-3) / 1e9 / pk.get("hbm_gbs"),
                "note": "pack + keys + radix index + join (count, scan, fill), timed with CUDA events inside the step; "
                        "includes one host round trip for the pair count; K6 is fused into the DP epilogue"}
        # ---- CPU baseline on this box's host cores (bounded sample)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            # a bounded random sample of the SAME candidate list the GPU just processed
            rs_, pa_, pb_, _, _ = device_step()
            n_s = min(pairs, 4_000_000)
            idx = torch.randperm(pairs, device=dev)[:n_s] if pairs > n_s else torch.arange(pairs, device=dev)
            pa_h, pb_h = pa_[idx].cpu().numpy(), pb_[idx].cpu().numpy()
            del rs_, pa_, pb_, idx
            cpu = cpu_arm(ub, uo, pa_h, pb_h, target_s=args.cpu_seconds, total_pairs=pairs)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "int16x2" if plan["mode"] == "packed16" else "int32",
                "data": "synthetic",
                "config": {"workload": args.workload, "k": args.k, "seed": args.seed, "reads": n_reads,
                           "unique_reads": U, "max_read_len": max_len, "candidate_pairs": pairs, "edges": n_edges,
                           "cells": cells, "edge_checksum": checksum, "sharding": f"pair-range x{world}", "exchange": exchange,
                           "l2": "flushed between timed iterations (256 MiB memset)",
                           "scoring": "match 10, mismatch -1, indel -2^31 (reference defaults)"},
                "pairs_per_s": pairs / (kmer_total_ms / args.steps * 1e-3),
                "stage_ms": {"kmer_index_join": kmer_total_ms / args.steps, "overlap_dp_fused_expand": dp_total_ms / args.steps,
                             "edge_count_scan_and_gather": ms_per_step - (kmer_total_ms + dp_total_ms) / args.steps},
                "kmer_stages": kmer,
                "e2e": {"value": cells / (e2e_per_step * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_per_step,
                        "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                        "checksum_matches_device_path": e2e_ok,
                        "path": "engine.overlap_edges: pinned host reads in, host edge rows out, D2H overlapped with the DP"
                                + ("; each rank writes its slice into one shared page-locked host buffer" if world > 1 else "")},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        sink.close()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # BASELINE.json quotes the metric "at 1/2/4/8 B200" on configs[2] (4.6 Mb genome, 1M reads, l=150,
    # p=0.005, sharded by read-ID range); it fits one GPU (~32 GB), so it is the workload at every N.
    # configs[1] is --workload phix_n50000_l150 (numbers in DESIGN.md).
    ap.add_argument("--workload", default="ecoli_n1m_l150")
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peer-stores", action="store_true", help="N>1: gather with NCCL instead of peer-memory stores")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
