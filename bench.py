#!/usr/bin/env python
"""bench.py -- edges/sec through the 40-iteration embed() loop (BASELINE.json's metric).

  python bench.py --gpus 1 --steps 2 --warmup 3                 # our arm, one B200
  torchrun ... bench.py --gpus N --steps K --warmup W           # row-sharded over N B200s (one rank per GPU)
  python bench.py --impl reference --gpus 1 --steps 2 --warmup 1  # the reference's CPU path (restated: oracle/)

A "step" is one full pass of the hot path over the workload: init -> 40 x (SpMM -> L2 -> whiten).  Default
workload = BASELINE.json configs[2], the configuration it lists for 1/2/4/8 GPUs: ogbn-products-shaped Chung-Lu graph,
2.45M nodes / 61.9M edges, d=256, iters=40 (it fits one GPU).  `--workload er-1m-20m-d256` is
configs[1] (C2), the round-1 default.
`value`  : E * iters / t with graph and state resident in HBM, timed with CUDA events on the library's stream.
`e2e`    : the same metric through the public host-buffer API (cleora_b200.embed on host arrays): CSR upload
           and result download inside the timed region.
`roofline`: SpMM (K1) algorithmic bytes / its mean launch time (CUDA events inside the library, same timed
           region) against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nodes, input edges, d, generator, rng seed)   -- SURVEY.md 8d
    "er-1m-20m-d256": dict(n=1_000_000, e=20_000_000, d=256, kind="er", seed=1),
    "products-2.4m-62m-d256": dict(n=2_449_029, e=61_859_140, d=256, kind="chunglu", seed=2, alpha=0.5),
    # BASELINE.json configs[4]: Twitter-2010-shaped, 41.65M nodes / 1.468B edges, Chung-Lu with degree
    # exponent ~2.2 (endpoint weight (i+10)^-0.833), d=128.  Too large for the numpy generator + host builder inside a
    # bench run: the pair stream is generated ON THE DEVICE (counter-based RNG, cleora_dev_synth_pairs) and the CSR is
    # built on the device (cleora_dev_graph_from_pairs); `cpu_sample` = the CPU arms run on a 1/16-scale graph of the
    # same family.  twitter-5m-183m-d128 is the 1/8-scale member used for quick checks.
    "twitter-41m-1.5b-d128": dict(n=41_652_230, e=1_468_365_182, d=128, kind="chunglu", seed=5, alpha=0.833,
                                  device_gen=True, cpu_sample=16),
    "twitter-5m-183m-d128": dict(n=5_206_528, e=183_545_648, d=128, kind="chunglu", seed=5, alpha=0.833,
                                 device_gen=True, cpu_sample=4),
    # BASELINE.json configs[3]: 10M hyperedge lines over a 1M-product vocabulary, one
    # `complex::reflexive::product` column (clique expansion, n = 1M), size 1 + Poisson(3) capped at 16 (no trimming),
    # members Zipf(1.0), d=512.  "edges" = hyperedge lines.  hyper-100k-1m-d512 is the 1/10-scale member.
    "hyper-1m-10m-d512": dict(n=1_000_000, e=10_000_000, d=512, kind="hyper", seed=3, cpu_sample=10),
    "hyper-100k-1m-d512": dict(n=100_000, e=1_000_000, d=512, kind="hyper", seed=3, cpu_sample=4),
    "er-200k-4m-d256": dict(n=200_000, e=4_000_000, d=256, kind="er", seed=1),
    "er-50k-1m-d128": dict(n=50_000, e=1_000_000, d=128, kind="er", seed=1),
}


def gen_pairs(w):
    """Synthetic edge list (numpy PCG64 default_rng(seed)); u != v, duplicates allowed (they merge)."""
    rs = np.random.default_rng(w["seed"])
    n, e = w["n"], w["e"]
    if w["kind"] == "er":
        u = rs.integers(0, n, size=e, dtype=np.int64)
        v = rs.integers(0, n, size=e, dtype=np.int64)
    else:  # Chung-Lu: endpoints drawn with probability ~ (i + i0)^-alpha
        i0 = 10.0
        wts = (np.arange(n, dtype=np.float64) + i0) ** (-w["alpha"])
        cdf = np.cumsum(wts)
        cdf /= cdf[-1]
        u = np.searchsorted(cdf, rs.random(e)).astype(np.int64)
        v = np.searchsorted(cdf, rs.random(e)).astype(np.int64)
        perm = rs.permutation(n)          # decouple node id from weight rank
        u, v = perm[u], perm[v]
    keep = u != v
    return u[keep].astype(np.uint32), v[keep].astype(np.uint32)


def gen_hyperedges(w, scale=1):
    """Hyperedge lines as (members uint32, offsets int64): sizes 1 + Poisson(3) capped at 16, members Zipf(1.0) over the
    vocabulary (inverse CDF), ids decoupled from the popularity rank by a random permutation."""
    rs = np.random.default_rng(w["seed"])
    vocab, lines = max(2, w["n"] // scale), max(1, w["e"] // scale)
    k = np.minimum(1 + rs.poisson(3.0, lines), 16)
    offsets = np.zeros(lines + 1, np.int64)
    np.cumsum(k, out=offsets[1:])
    cdf = np.cumsum(1.0 / np.arange(1, vocab + 1))
    cdf /= cdf[-1]
    perm = rs.permutation(vocab)
    members = perm[np.searchsorted(cdf, rs.random(int(offsets[-1])))].astype(np.uint32)
    return members, offsets


def build_host_graph(w):
    """The product's host-side graph for a host-built workload -> (SparseMatrix, input edge count E)."""
    import cleora_b200 as cb
    if w["kind"] == "hyper":
        members, offsets = gen_hyperedges(w)
        return cb.SparseMatrix.from_hyperedge_arrays(members, offsets, "complex::reflexive::product"), int(len(offsets) - 1)
    u, v = gen_pairs(w)
    return cb.SparseMatrix.from_edge_arrays(u, v), int(len(u))


def build_oracle_graph(w):
    """The CPU arms' graph, built by the oracle alone -> (OracleGraph, edge count, description of the sample)."""
    import oracle
    scale = w.get("cpu_sample", 1) if (w.get("device_gen") or w["kind"] == "hyper") else 1
    if w["kind"] == "hyper":
        members, offsets = gen_hyperedges(w, scale)
        lines = [" ".join(map(str, members[offsets[i]:offsets[i + 1]])) for i in range(len(offsets) - 1)]
        og, e = oracle.build_graph(lines, "complex::reflexive::product"), len(lines)
    else:
        u, v = gen_pairs_continuous(w, scale) if w.get("device_gen") else gen_pairs(w)
        og, e = oracle.graph_from_pairs(u, v), len(u)
    what = ("the full workload" if scale == 1 else
            f"a 1/{scale}-scale graph of the same family ({og.n} nodes / {e} edges; edges/s is size-normalised)")
    return og, int(e), what


def gen_pairs_continuous(w, scale=1):
    """Host restatement of the device generator's distribution for `device_gen` workloads (same family, not the same
    stream): n / scale nodes, e / scale pairs, endpoint rank from the continuous inverse CDF of (x + 10)^-alpha, ids
    decoupled from the rank by a random permutation.  Used by the CPU arms only."""
    rs = np.random.default_rng(w["seed"])
    n, e = max(2, w["n"] // scale), max(1, w["e"] // scale)
    if w["kind"] == "er":
        u, v = rs.integers(0, n, size=e, dtype=np.int64), rs.integers(0, n, size=e, dtype=np.int64)
    else:
        e1, i0 = 1.0 - w["alpha"], 10.0
        a, b = i0 ** e1, (n + i0) ** e1
        perm = rs.permutation(n)
        draw = lambda: perm[np.minimum(((a + rs.random(e) * (b - a)) ** (1.0 / e1) - i0).astype(np.int64), n - 1)]  # noqa: E731
        u, v = draw(), draw()
    keep = u != v
    return u[keep].astype(np.uint32), v[keep].astype(np.uint32)


def spmm_bytes(n, nnz, d):
    """B_spmm = nnz*(4+4+4d) + 8(n+1) + 4nd  (no-reuse gather model)."""
    return nnz * (8 + 4 * d) + 8 * (n + 1) + 4 * n * d


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference (CPU) arm
def cpu_loop(og, d, iters, whiten, threads):
    """The reference's CPU path restated (oracle/): Rust SpMM order + numpy L2 + numpy whitening."""
    import oracle
    x = oracle.init_matrix(og.hashes, d, 0)
    t0 = time.perf_counter()
    if whiten:
        for _ in range(iters):
            x = oracle.whiten_embeddings(oracle.normalize(oracle.spmm(og, x), "l2"))
    else:
        x = oracle.embed_fast(og, d, iters, x0=x)
    return time.perf_counter() - t0


def oracle_graph_from(g):
    import oracle
    rowptr, col, left, sym = g._csr()
    return oracle.OracleGraph(rowptr, col, left, sym, g.entity_degrees, g.entity_hashes(),
                              np.zeros(g.num_entities, np.uint8), None)


def run_reference(args, w, name):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crate cannot be built
    here (no cargo/rustc), so this is the oracle port: C/OpenMP restatement of src/embedding.rs + the
    reference's numpy whitening, all host threads.  One step = `sample_iters` iterations of the full workload.
    Nothing of the product is loaded in this arm: the CSR comes from the oracle's own integer ingest."""
    import oracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    og, n_edges, what = build_oracle_graph(w)           # host-only CSR construction, outside the timed region
    sample_iters = args.cpu_iters
    for _ in range(min(args.warmup, 1)):
        cpu_loop(og, w["d"], 1, args.whiten, cores)
    times = [cpu_loop(og, w["d"], sample_iters, args.whiten, cores) for _ in range(args.steps)]
    t = sum(times) / len(times)
    value = n_edges * sample_iters / t
    line = {
        "impl": "reference", "metric": "edges/sec through the 40-iteration embed() loop", "value": value,
        "unit": "edges/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "nodes": og.n, "edges": n_edges, "nnz": og.nnz, "d": w["d"],
                   "iters": args.iters, "whiten": bool(args.whiten)},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_iters} of {args.iters} iterations per step of {what}"
                                   " (restated Rust/rayon SpMM+L2 in C/OpenMP + the reference's numpy whitening)"},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "product_library_loaded": any("libcleora_b200" in ln for ln in open("/proc/self/maps")),
    }
    print(json.dumps(line), flush=True)


def torch_baseline(g, d, sample_iters):
    """Second stated baseline: the reference's OWN GPU path, `propagate_gpu`
    (pycleora/__init__.py:684-739: coalesced COO `torch.sparse.mm` (cuSPARSE), row L2 norm, and
    `_whiten_embeddings_torch`, :979-997: f32 mean / covariance GEMM, `torch.linalg.eigh`, rsqrt-scaled transform
    GEMM), restated here stage for stage with the same torch calls, on the same GPU, same graph.  Returns seconds per
    iteration (CUDA events; one warm-up iteration)."""
    import torch
    rows, cols, vals, n, _ = g.to_sparse_csr("left")
    idx = torch.stack([torch.from_numpy(rows.astype(np.int64)), torch.from_numpy(cols.astype(np.int64))])
    adj = torch.sparse_coo_tensor(idx, torch.from_numpy(vals), size=(n, n)).to("cuda").coalesce()
    del idx
    emb = torch.rand((n, d), device="cuda", dtype=torch.float32) * 2 - 1

    def one(emb):
        emb = torch.sparse.mm(adj, emb)
        emb = emb / torch.norm(emb, dim=1, keepdim=True).clamp(min=1e-10)
        mean = emb.mean(dim=0, keepdim=True)
        centered = emb - mean
        cov = centered.transpose(0, 1).matmul(centered) / max(n - 1, 1)
        w_, v_ = torch.linalg.eigh(cov)
        order = torch.argsort(w_, descending=True)
        transform = v_[:, order] * torch.rsqrt(torch.clamp(w_[order], min=1e-10)).unsqueeze(0)
        return centered.matmul(transform)

    emb = one(emb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(sample_iters):
        emb = one(emb)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / sample_iters


def parity_probe(d, iters):
    """Small in-process parity check of the configuration being timed (not in the timed region): ER 20k nodes /
    200k pairs, the workload's d and iteration count, default options, against the CPU oracle's reference-order
    loop (pycleora/__init__.py:109-125,963-971).  Bars as in tests/test_gpu_parity.py."""
    import cleora_b200 as cb
    import oracle
    from tests.helpers import gram_err, procrustes_err
    rs = np.random.default_rng(7)
    n, e = 20000, 200000
    u, v = rs.integers(0, n, e), rs.integers(0, n, e)
    k = u != v
    g = cb.SparseMatrix.from_edge_arrays(u[k], v[k])
    og = oracle.graph_from_pairs(u[k], v[k])
    t0 = time.perf_counter()
    got = cb.embed(g, d, iters)
    ref = oracle.embed(og, d, iters)
    pe, ge = procrustes_err(got, ref), gram_err(got, ref)
    fast = cb.embed(g, d, iters, whiten=False)
    fref = oracle.embed(og, d, iters, whiten=False)
    fe = float(np.max(np.abs(fast.astype(np.float64) - fref)) / np.max(np.abs(fref)))
    return {"graph": f"ER {n} nodes / {int(k.sum())} pairs, d={d}, {iters} iterations, default options",
            "whiten_procrustes_err": pe, "whiten_gram_err": ge, "whiten_ok": bool(pe <= 1e-4 and ge <= 1e-5),
            "nowhiten_max_err_of_scale": fe, "nowhiten_ok": bool(fe <= 1e-5),
            "seconds": round(time.perf_counter() - t0, 2)}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, w, name):
    import torch
    import cleora_b200 as cb
    from cleora_b200 import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    L = _lib.lib()
    _lib.check(L.cleora_set_device(local))
    if world > 1:
        from cleora_b200 import sharded
        return sharded.bench(args, w, name, build_host_graph, spmm_bytes, measured_peaks, ClockSampler)

    d, iters = w["d"], args.iters
    device_gen = bool(w.get("device_gen"))
    if device_gen:          # pair stream and CSR made on the device (see WORKLOADS); nothing of the graph on the host
        u_d, v_d = cb.synth_pairs(w["kind"], w["n"], w["e"], w["seed"], w.get("alpha", 0.5))
        g = cb.SparseMatrix.from_edge_arrays_device(u_d, v_d, want_sym=False)
        E = int(w["e"])
        u = v = None
    else:
        g, E = build_host_graph(w)
    n, nnz = g.num_entities, g.num_edges
    norm = _lib.NORM_L2_NUMPY if args.whiten else _lib.NORM_L2_RUST
    dev_out = torch.empty((n, d), dtype=torch.float32, device="cuda")     # result stays in HBM for `value`
    timings = np.zeros(8)

    def step_resident(t=None):
        done = np.zeros(1, np.int64)
        if args.whiten:
            _lib.check(L.cleora_embed(g._handle(), None, d, iters, 0, 0, 0.0, 0.0, norm, 1, _lib.f32p(dev_out.data_ptr()),
                                      done.ctypes.data_as(_lib.c_i64p),
                                      None if t is None else t.ctypes.data_as(_lib.c_f64p)))
        else:
            _lib.check(L.cleora_embed(g._handle(), None, d, iters, 0, 0, 0.0, 0.0, norm, 0, _lib.f32p(dev_out.data_ptr()),
                                      done.ctypes.data_as(_lib.c_i64p),
                                      None if t is None else t.ctypes.data_as(_lib.c_f64p)))

    _lib.check(L.cleora_dev_graph_prepare(g._handle()))                 # CSR resident before the timed region
    host_eigh = _lib.auto_host_eigh(n, d, iters, norm, bool(args.whiten), 0.0, 0.0)   # what cleora_b200.embed() would pick
    eigh_ctx = _lib.host_eigh(host_eigh)
    eigh_ctx.__enter__()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = L.cleora_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()                                                          # legacy default stream == library stream
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        t_step = np.zeros(8)
        step_resident(t_step)                  # the library reports per-call phase totals (CUDA events, stream 0)
        timings += t_step
    ev1.record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall
    launches = L.cleora_kernel_launch_count() - launches0
    clk = clocks.stop()
    ms_step = ev0.elapsed_time(ev1) / args.steps
    value = E * iters / (ms_step * 1e-3)

    # ---- e2e: public host API on HOST buffers, host<->device copies inside the timed region.
    e2e_times = []
    if not device_gen:
        # every step copies the CSR host -> device again (same device buffers, pinned host arrays) and the result
        # device -> pinned host memory.  X0 is produced on the device by the init kernel from the uploaded entity
        # hashes (the reference also initialises inside embed()).
        pinned = cb.pinned_empty((n, d), np.float32)
        for i in range(1 + args.e2e_steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _lib.check(L.cleora_graph_refresh_device(g._handle(), None))
            g.embed_device(d, iters, "left", norm, 0, None, 0.0, 0.0, bool(args.whiten), out=pinned)
            torch.cuda.synchronize()
            if i > 0:
                e2e_times.append(time.perf_counter() - t0)
        h2d = 8 * (n + 1) + nnz * (4 + 4) + 8 * n     # rowptr + col + left values + hashes (sym values only on first symmetric use)
        e2e_note = ("cleora_b200 host API on host buffers: CSR re-copied host->device and result copied to pinned host "
                    "memory every step; X0 comes from the init kernel (entity hashes are part of the upload)")
    else:
        # the edge list lives in pinned host memory; every step uploads it, builds the CSR on the device
        # (from_edge_arrays_device), runs the loop and copies the result to pinned host memory
        hu, hv = cb.pinned_empty((E,), np.int32), cb.pinned_empty((E,), np.int32)
        torch.from_numpy(hu).copy_(u_d)
        torch.from_numpy(hv).copy_(v_d)
        del g, dev_out
        L.cleora_release_workspace()
        torch.cuda.empty_cache()
        pinned = cb.pinned_empty((n, d), np.float32)
        for i in range(1 + max(1, args.e2e_steps - 1)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            u_d.copy_(torch.from_numpy(hu), non_blocking=True)
            v_d.copy_(torch.from_numpy(hv), non_blocking=True)
            g2 = cb.SparseMatrix.from_edge_arrays_device(u_d, v_d, want_sym=False)
            g2.embed_device(d, iters, "left", norm, 0, None, 0.0, 0.0, bool(args.whiten), out=pinned)
            torch.cuda.synchronize()
            if i > 0:
                e2e_times.append(time.perf_counter() - t0)
            del g2
        g = None
        h2d = 8 * E
        e2e_note = ("edge list (2 x uint32 per edge) in pinned host memory -> device, CSR built on the device "
                    "(cleora_dev_graph_from_pairs), loop, result copied to pinned host memory -- all inside the timed region")
    eigh_ctx.__exit__(None, None, None)
    e2e_t = sum(e2e_times) / len(e2e_times)
    d2h = 4 * n * d

    # ---- roofline of the dominant kernel (K1 SpMM+L2), live CUDA-event time from the same timed region
    peak, peak_src = measured_peaks()
    spmm_ms = timings[2] / (iters * args.steps)
    b_spmm = spmm_bytes(n, nnz, d)
    achieved = b_spmm / (spmm_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "spmm_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(name)

    # ---- parity of the timed configuration (small graph, against the CPU oracle; outside the timed region)
    parity = None
    if not args.no_parity:
        parity = parity_probe(d, iters)

    # ---- second stated baseline: the reference's own torch GPU path on this GPU (bounded sample)
    tbase = None
    if args.torch_iters > 0 and device_gen:
        tbase = {"unavailable": "graph is device-resident only (no COO on the host); run a host-built workload"}
    elif args.torch_iters > 0:
        try:
            L.cleora_release_workspace()
            torch.cuda.empty_cache()
            tt = torch_baseline(g, d, args.torch_iters)
            tbase = {"value": E / tt, "unit": "edges/s", "ms_per_iter": 1e3 * tt,
                     "what": "reference's propagate_gpu restated with the same torch calls (pycleora/__init__.py:684-739,"
                             "979-997): coalesced COO torch.sparse.mm + row norm + f32 covariance GEMM + torch.linalg.eigh "
                             "+ transform GEMM", "sample": f"{args.torch_iters} iterations after 1 warm-up"}
        except Exception as ex:  # noqa: BLE001 - a baseline must not break the bench line
            tbase = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}

    # ---- CPU baseline beside it: bounded sample of the same workload on the host cores
    cpu = None
    if not args.no_cpu_baseline:
        import oracle
        cores = os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        if device_gen or w["kind"] == "hyper":
            og, e_cpu, what = build_oracle_graph(w)
        else:
            og, e_cpu, what = oracle_graph_from(g), E, "the full workload"
        cpu_loop(og, d, 1, args.whiten, cores)
        tc = cpu_loop(og, d, args.cpu_iters, args.whiten, cores)
        cpu = {"value": e_cpu * args.cpu_iters / tc, "unit": "edges/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_iters} of {iters} iterations of {what}, restated Rust/rayon SpMM+L2 "
                         "(C/OpenMP, oracle/) + the reference's numpy whitening"}

    chol = bool(L.cleora_get_option(b"chol_whiten")) and bool(args.whiten)
    line = {
        "metric": "edges/sec through the 40-iteration embed() loop", "value": value, "unit": "edges/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic" + (" (generated and ingested on the device)" if device_gen else ""),
        "config": {"workload": name, "nodes": n, "edges": E, "nnz": nnz, "d": d, "iters": iters,
                   "whiten": bool(args.whiten),
                   "inner_whitening": ("Cholesky factor on the device (iterations 0..T-2), PCA eigh on the last" if chol
                                       else "PCA eigh every iteration"),
                   "eigh": "numpy LAPACK on the host" if host_eigh else "cuSOLVER Dsyevd",
                   "pipeline_whiten": bool(L.cleora_get_option(b"pipeline_whiten")),
                   "l2_flush": "inputs (X + CSR per iteration) exceed the 126 MB L2"},
        "nnz_per_s": nnz * iters / (ms_step * 1e-3),
        "e2e": {"value": E * iters / e2e_t, "unit": "edges/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_t, "samples": len(e2e_times), "ms_min": 1e3 * min(e2e_times), "note": e2e_note},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "spmm_rows_kernel (K1: SpMM + fused L2)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": b_spmm, "ms_per_launch": spmm_ms},
        "phase_ms_per_iter": {k: timings[i] / (iters * args.steps) for i, k in
                              enumerate(["h2d", "init", "spmm", "stats", "eigh", "apply", "rmse", "d2h"])},
        "wall_ms_per_step": 1e3 * t_wall / args.steps,
        "parity": parity,
        "torch_baseline": tbase,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def main():
    # torchrun exports OMP_NUM_THREADS=1 for every rank.  The host-side CSR construction (rank 0) and the CPU arms are
    # OpenMP / OpenBLAS code, so give them the cores back -- all of them for the reference arm (only rank 0 works
    # there), an even share per rank otherwise.  Must happen before the libraries initialise their thread pools.
    if "WORLD_SIZE" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
        cores = os.cpu_count() or 1
        world = max(1, int(os.environ.get("WORLD_SIZE", "1")))
        share = cores if "reference" in sys.argv else max(1, cores // world)
        os.environ["OMP_NUM_THREADS"] = str(share)
        os.environ.setdefault("OPENBLAS_NUM_THREADS", str(min(share, 64)))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="products-2.4m-62m-d256", choices=sorted(WORKLOADS))
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--whiten", type=int, default=1)
    ap.add_argument("--cpu-iters", type=int, default=1, help="iterations per CPU sample (bounded baseline)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--torch-iters", type=int, default=3, help="iterations of the reference's torch GPU path (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, args.workload)
    else:
        run_ours(args, w, args.workload)


if __name__ == "__main__":
    main()
