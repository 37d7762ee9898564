#!/usr/bin/env python
"""Stand-alone time of the Cholesky whitening kernel (chol_whiten.cu) and of cuSOLVER's Dsyevd path for the same
covariance, d in {128, 256, 512}: CUDA events around 50 back-to-back launches."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleora_b200 import _lib  # noqa: E402

L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
for d in (128, 256, 512):
    rs = np.random.default_rng(d)
    a = rs.standard_normal((4 * d, d)) * rs.uniform(0.5, 2.0, d)
    cov = torch.from_numpy(np.cov(a, rowvar=False)).cuda()
    T = torch.empty(d, d, dtype=torch.float32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    res = {}
    for name, fn in (("chol", lambda: _lib.check(L.cleora_dev_chol_whiten(cov.data_ptr(), d, T.data_ptr(), status.data_ptr(), st))),
                     ("eigh", lambda: _lib.check(L.cleora_dev_whiten_transform(cov.data_ptr(), d, d, T.data_ptr(), st)))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 50
    print(f"d={d}: chol_whiten {res['chol']*1e3:.0f} us, cuSOLVER Dsyevd + build_transform {res['eigh']*1e3:.0f} us, status {int(status.item())}", flush=True)
