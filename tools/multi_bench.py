#!/usr/bin/env python
"""In-process multi-GPU embed (cleora_embed_multi) against the one-GPU call on the same graph, through the public host
API: CSR resident on the devices, result written to pinned host memory, wall clock around the call (so the numbers are
end-to-end-style: they include the result download and, for the multi-GPU call, thread start-up and buffer allocation).
    python tools/multi_bench.py --devices 0,1 [--workload er-1m-20m-d256] [--iters 40] [--reps 3]
Also prints the maximal element difference / Procrustes error of the two results (whitened loop: defined up to rotation)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import cleora_b200 as cb  # noqa: E402
from cleora_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="0,1")
    ap.add_argument("--workload", default="er-1m-20m-d256")
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--whiten", type=int, default=1)
    args = ap.parse_args()
    devices = [int(v) for v in args.devices.split(",")]
    w = bench.WORKLOADS[args.workload]
    d = w["d"]
    g, E = bench.build_host_graph(w)
    n = g.num_entities
    out = cb.pinned_empty((n, d), np.float32)
    norm = _lib.NORM_L2_NUMPY

    def run(devs):
        cb.set_devices(devs)
        try:
            ts = []
            for _ in range(args.reps + 1):                    # first repetition uploads the CSR / warms the allocator
                t0 = time.perf_counter()
                g.embed_device(d, args.iters, "left", norm, 0, None, 0.0, 0.0, bool(args.whiten), out=out)
                ts.append(time.perf_counter() - t0)
            return min(ts[1:]), out.copy()
        finally:
            cb.set_devices(None)

    t1, r1 = run(None)
    print(f"{args.workload}: 1 GPU   {1e3 * t1:9.1f} ms/call  {E * args.iters / t1 / 1e9:.3f} G edges/s (host API, result to pinned host)", flush=True)
    tn, rn = run(devices)
    print(f"{args.workload}: devices {devices}  {1e3 * tn:9.1f} ms/call  {E * args.iters / tn / 1e9:.3f} G edges/s  speed-up {t1 / tn:.2f}x", flush=True)
    m = min(n, 200000)
    a, b = rn[:m].astype(np.float64), r1[:m].astype(np.float64)
    if args.whiten:
        u, _, vt = np.linalg.svd(a.T @ b)
        err = float(np.max(np.abs(a @ (u @ vt) - b)) / np.max(np.abs(b)))
        print(f"Procrustes error multi vs single (first {m} rows): {err:.2e}", flush=True)
    else:
        print(f"bit-identical to one GPU: {bool(np.array_equal(rn, r1))}", flush=True)


if __name__ == "__main__":
    main()
