#!/usr/bin/env python
"""K1 locality probe (VERDICT r1 "next" #7): does running the SpMM column slice by column slice -- so that the gathered
slice X[:, g*ds:(g+1)*ds] (n x ds floats) is small enough to live in the 126 MB L2 -- beat one pass over full rows?

One GPU plays all G slices in turn with the kernels of the column-sharded loop (cleora_dev_spmm_scatter, one owner):
    for g in range(G):  W[:, g*ds:(g+1)*ds] = A @ Xs[g]        (Xs[g] contiguous [n, ds])
against the production kernel W = A @ X on row-major X.  Prints ms per full product for G in {1 (baseline), 2, 4, 8, 16}.
Usage: python tools/k1_slices_probe.py [workload]   (bench.py workload names; default er-1m-20m-d256)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cleora_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "er-1m-20m-d256"
w = bench.WORKLOADS[name]
g, E = bench.build_host_graph(w)
L = _lib.lib()
n, d, nnz = g.num_entities, w["d"], g.num_edges
_lib.check(L.cleora_dev_graph_prepare(g._handle()))
st = torch.cuda.current_stream().cuda_stream
x = torch.randn(n, d, device="cuda")
out = torch.empty(n, d, device="cuda")


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


base = timed(lambda: _lib.check(L.cleora_dev_spmm(g._handle(), 0, x.data_ptr(), d, out.data_ptr(), None, 1.0, 0.0, 0, st)))
print(f"{name}: n={n} nnz={nnz} d={d}")
print(f"G=1 (row-major, production K1): {base:.3f} ms")
ref = out.clone()
for G in (2, 4, 8, 16, 32):
    ds = d // G
    if ds not in (8, 16, 32, 64, 96, 128, 192, 256, 384, 512):
        continue
    xs = [x[:, i * ds:(i + 1) * ds].contiguous() for i in range(G)]
    dst = (C.c_void_p * 1)(out.data_ptr())
    out.zero_()

    def run():
        for i in range(G):
            _lib.check(L.cleora_dev_spmm_scatter(g._handle(), 0, xs[i].data_ptr(), ds, dst, 1, n, d, i * ds, None, 1.0, 0.0, st))

    t = timed(run)
    same = torch.equal(out, ref)
    print(f"G={G} (slices of {ds} floats = {n * ds * 4 / 1e6:.0f} MB each): {t:.3f} ms  ({base / t:.2f}x)  bit-identical: {same}")
