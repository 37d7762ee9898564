#!/usr/bin/env python
"""K3 (whiten_apply_tc_kernel) A/B: (BK 32, 2 stages) vs the same with coalesced producers and SWIZZLE_128B A tiles vs
(BK 16, 4 stages), on the variants the loop
launches -- plain apply, fused row norm + row scale (pipelined loop), and the same with an upper-triangular transform
(Cholesky inner iterations).  Prints time per launch and the error against an f64 reference of the first rows; the two
shapes must agree bit for bit (same products in the same order per accumulator column)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from cleora_b200 import _lib  # noqa: E402

L = _lib.lib()
# (name, K floats per stage, A-tile layout): round-1 shape | coalesced producers + SWIZZLE_128B A tiles | 4 half-size stages
VARIANTS = [("bk32", 32, 0), ("bk32+asw", 32, 1), ("bk16x4", 16, 0)]
if len(sys.argv) > 1:
    VARIANTS = [v for v in VARIANTS if v[0] in sys.argv[1].split(",")]


def run_case(n, d, fused, upper, reps=8):
    rs = np.random.default_rng(n + d)
    x = rs.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    mean = x.mean(axis=0).astype(np.float32)
    T = (rs.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32)
    if upper:
        T = np.triu(T)
    rowscale = (0.5 + rs.random(n)).astype(np.float32)
    xd, md, Td, rd = (torch.from_numpy(a).cuda() for a in (x, mean, np.ascontiguousarray(T), rowscale))
    st = torch.cuda.current_stream().cuda_stream
    outs = {}
    for name, bk, asw in VARIANTS:
        _lib.check(L.cleora_set_option(b"k3_bk", bk))
        _lib.check(L.cleora_set_option(b"k3_asw", asw))
        out = torch.empty(n, d, dtype=torch.float32, device="cuda")

        def go():
            if fused:
                _lib.check(L.cleora_dev_whiten_apply_ex(xd.data_ptr(), n, d, md.data_ptr(), Td.data_ptr(), d, out.data_ptr(),
                                                        _lib.NORM_L2_NUMPY, rd.data_ptr(), 1 if upper else 0, st))
            else:
                _lib.check(L.cleora_dev_whiten_apply(xd.data_ptr(), n, d, md.data_ptr(), Td.data_ptr(), d, out.data_ptr(), st))
        go()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            go()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        m = 4096
        a = x[:m].astype(np.float64) - (rowscale[:m, None].astype(np.float64) if fused else 1.0) * mean.astype(np.float64)
        ref = a @ T.astype(np.float64)
        if fused:
            ref /= np.maximum(np.linalg.norm(ref, axis=1, keepdims=True), 1e-10)
        got = out[:m].cpu().numpy()
        err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
        outs[name] = out
        print(f"n={n} d={d} fused={int(fused)} upper={int(upper)} {name}: {ms:.3f} ms  err/scale {err:.2e}", flush=True)
    same = all(bool(torch.equal(outs[VARIANTS[0][0]], outs[v[0]])) for v in VARIANTS[1:])
    print(f"    all variants identical to {VARIANTS[0][0]}: {same}", flush=True)
    return same


if __name__ == "__main__":
    ok = True
    for n, d in [(1000000, 256), (2449029, 256), (1000000, 128), (300000, 512), (5000, 64)]:
        for fused, upper in [(False, False), (True, False), (True, True)]:
            ok &= run_case(n, d, fused, upper)
    _lib.check(L.cleora_set_option(b"k3_bk", 32))
    _lib.check(L.cleora_set_option(b"k3_asw", 0))
    print("ALL IDENTICAL" if ok else "MISMATCH")
