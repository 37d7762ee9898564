#!/usr/bin/env python
"""tcgen05 K3 (whiten_apply_tc_kernel) vs the SIMT kernel vs the oracle: accuracy and time, several shapes."""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import oracle
    from cleora_b200 import _lib
    L = _lib.lib()
    mode = os.environ.get("CLEORA_B200_APPLY", "tc")
    for n, d in [(1000, 32), (4097, 64), (5000, 256), (300000, 256), (1000000, 256), (200000, 128)]:
        rs = np.random.default_rng(d + n)
        x = oracle.normalize(rs.standard_normal((n, d)).astype(np.float32) + 0.05)
        mean, cov = oracle.whiten_stats(x)
        T = oracle.whiten_transform(cov)
        ref = oracle.whiten_apply(x[:20000], mean, T)
        xd = torch.from_numpy(x).cuda()
        md = torch.from_numpy(mean.astype(np.float32)).cuda()
        Td = torch.from_numpy(np.ascontiguousarray(T)).cuda()
        out = torch.empty(n, d, dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        def run():
            _lib.check(L.cleora_dev_whiten_apply(xd.data_ptr(), n, d, md.data_ptr(), Td.data_ptr(), d, out.data_ptr(), st))
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record(); torch.cuda.synchronize()
        got = out[:20000].cpu().numpy()
        err = float(np.max(np.abs(got.astype(np.float64) - ref)) / np.max(np.abs(ref)))
        tail = out[-1].cpu().numpy()
        ref_tail = (x[-1] - mean.astype(np.float32)) @ T
        err_tail = float(np.max(np.abs(tail - ref_tail)) / np.max(np.abs(ref_tail)))
        print(f"[{mode}] n={n} d={d}: {e0.elapsed_time(e1) / 5:.3f} ms, err/scale {err:.2e}, last-row err {err_tail:.2e}", flush=True)
else:
    for mode in ("tc", "simt"):
        env = dict(os.environ, CLEORA_B200_APPLY=mode)
        try:
            r = subprocess.run([sys.executable, __file__, "child"], env=env, timeout=240, capture_output=True, text=True)
            print(r.stdout[-3000:], r.stderr[-1500:], flush=True)
        except subprocess.TimeoutExpired as e:
            print(f"[{mode}] TIMEOUT", (e.stdout or b"")[-2000:], flush=True)
