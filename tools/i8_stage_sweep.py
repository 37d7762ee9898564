#!/usr/bin/env python
"""Sweep the int8 Gram kernel's raw/plane stage counts (n=1M, d=256); each configuration in its own process."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from cleora_b200 import _lib
    L = _lib.lib()
    n, d = 1_000_000, 256
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.nn.functional.normalize(torch.randn(n, d, device="cuda", generator=g) + 0.1, dim=1).contiguous()
    sums = torch.zeros(d, dtype=torch.float64, device="cuda")
    cov = torch.zeros(d, d, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.cleora_dev_col_sums(x.data_ptr(), n, d, sums.data_ptr(), 0, st))
    mean = sums / n
    run = lambda: _lib.check(L.cleora_dev_centered_gram(x.data_ptr(), n, d, mean.data_ptr(), cov.data_ptr(), st))
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record(); torch.cuda.synchronize()
    print(f"RS={os.environ.get('CLEORA_B200_I8_RAW_STAGES')} PS={os.environ.get('CLEORA_B200_I8_PLANE_STAGES')}: "
          f"{e0.elapsed_time(e1) / 10:.3f} ms  trace {float(torch.trace(cov)) / (n - 1):.6f}", flush=True)
else:
    for rs, ps in ((3, 3), (4, 2), (5, 1), (2, 2), (1, 1), (2, 4)):
        env = dict(os.environ, CLEORA_B200_I8_RAW_STAGES=str(rs), CLEORA_B200_I8_PLANE_STAGES=str(ps))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, timeout=120, capture_output=True, text=True)
        print(r.stdout.strip()[-300:], r.stderr.strip()[-300:], flush=True)
