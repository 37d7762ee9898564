#!/usr/bin/env python
"""K2b variants (tcgen05 int8-exact vs FP64 DMMA): agreement and time at bench shapes, data generated on the GPU."""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SHAPES = [(1_000_000, 256), (300_000, 128), (1_000_000, 512), (500_000, 384)]

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from cleora_b200 import _lib
    L = _lib.lib()
    mode = os.environ.get("CLEORA_B200_GRAM", "i8")
    out = {}
    for n, d in SHAPES:
        g = torch.Generator(device="cuda").manual_seed(n + d)
        x = torch.randn(n, d, device="cuda", generator=g) * torch.linspace(0.3, 2.0, d, device="cuda") + 0.1
        x = torch.nn.functional.normalize(x, dim=1).contiguous()
        sums = torch.zeros(d, dtype=torch.float64, device="cuda")
        cov = torch.zeros(d, d, dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(L.cleora_dev_col_sums(x.data_ptr(), n, d, sums.data_ptr(), 0, st))
        mean = sums / n

        def run():
            _lib.check(L.cleora_dev_centered_gram(x.data_ptr(), n, d, mean.data_ptr(), cov.data_ptr(), st))
        results = {}
        for needed in (((0, 1) if d <= 256 else (1,)) if mode == "i8" else (0,)):   # int8: all columns staged vs the needed ones (d > 256: compact only)
            _lib.check(L.cleora_set_option(b"gram_needed_cols", needed))
            cov.zero_()
            run(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                run()
            e1.record(); torch.cuda.synchronize()
            results[needed] = cov.cpu().numpy().copy()
            if mode == "i8":
                print(f"[i8 needed_cols={needed}] n={n} d={d}: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
        if mode == "i8" and d <= 256:
            print(f"[i8] n={n} d={d}: needed-columns result identical to all-columns: "
                  f"{bool(np.array_equal(results[0], results[1]))}", flush=True)
        c = results[max(results)]
        ref = torch.cov(x[:200000].double().T).cpu().numpy() if n >= 200000 else None
        np.save(f"/tmp/gram_{mode}_{n}_{d}.npy", c)
        sym = float(np.max(np.abs(c - c.T)) / np.max(np.abs(c)))
        print(f"[{mode}] n={n} d={d}: {e0.elapsed_time(e1) / 5:.3f} ms, asym {sym:.1e}, trace/n {np.trace(c) / (n - 1):.6f}", flush=True)
else:
    for mode in os.environ.get("GRAM_CHECK_MODES", "i8,v3").split(","):
        env = dict(os.environ, CLEORA_B200_GRAM=mode)
        try:
            r = subprocess.run([sys.executable, __file__, "child"], env=env, timeout=300, capture_output=True, text=True)
            print(r.stdout[-2500:], r.stderr[-2500:], flush=True)
        except subprocess.TimeoutExpired as e:
            print(f"[{mode}] TIMEOUT", (e.stdout or b"")[-2000:], flush=True)
    for n, d in SHAPES:
        try:
            a, b = np.load(f"/tmp/gram_i8_{n}_{d}.npy"), np.load(f"/tmp/gram_v3_{n}_{d}.npy")
            print(f"i8 vs DMMA n={n} d={d}: max|diff|/max|cov| = {np.max(np.abs(a - b)) / np.max(np.abs(b)):.3e}", flush=True)
        except Exception as e:  # noqa: BLE001
            print("compare failed", e)
