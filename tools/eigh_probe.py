#!/usr/bin/env python
"""Times the d x d symmetric eigensolver candidates on this box (host LAPACK vs cuSOLVER vs torch)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from threadpoolctl import ThreadpoolController  # noqa: E402

ctl = ThreadpoolController()
for d in (128, 256):
    c = np.cov(np.random.default_rng(0).standard_normal((20000, d)), rowvar=False)
    for lim in (1, 4):
        def run():
            if lim is None:
                return np.linalg.eigh(c)
            with ctl.limit(limits=lim, user_api="blas"):
                return np.linalg.eigh(c)
        run()
        t = time.perf_counter()
        for _ in range(5):
            run()
        print(f"d={d} numpy eigh threads={lim}: {(time.perf_counter() - t) / 5 * 1e3:.2f} ms", flush=True)
    try:
        from scipy.linalg.lapack import dsyevd, dsyevr
        for name, fn in (("dsyevd", dsyevd), ("dsyevr", dsyevr)):
            for lim in (1, 4):
                with ctl.limit(limits=lim, user_api="blas"):
                    fn(c, lower=1)
                    t = time.perf_counter()
                    for _ in range(5):
                        fn(c, lower=1)
                print(f"d={d} scipy {name} threads={lim}: {(time.perf_counter() - t) / 5 * 1e3:.2f} ms", flush=True)
    except Exception as e:  # noqa: BLE001
        print("scipy probe failed:", e)
    try:
        import ctypes as C
        from cleora_b200 import _lib
        os.environ["CLEORA_B200_EIGH"] = "cusolver"
        L = _lib.lib()
        L.cleora_set_eigh(C.cast(None, _lib.EIGH_FN), None)
        T = np.empty((d, d), np.float32)
        for i in range(6):
            if i == 1:
                t = time.perf_counter()
            _lib.check(L.cleora_whiten_transform_from_cov(c.ctypes.data_as(_lib.c_f64p), d, d, T.ctypes.data_as(_lib.c_f32p)))
        print(f"d={d} cuSOLVER Dsyevd (incl. copies): {(time.perf_counter() - t) / 5 * 1e3:.2f} ms", flush=True)
        import torch
        ct = torch.from_numpy(c).cuda()
        torch.linalg.eigh(ct); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(5):
            torch.linalg.eigh(ct)
        torch.cuda.synchronize()
        print(f"d={d} torch.linalg.eigh cuda f64: {(time.perf_counter() - t) / 5 * 1e3:.2f} ms", flush=True)
    except Exception as e:  # noqa: BLE001
        print("gpu eigh probe failed:", e)
