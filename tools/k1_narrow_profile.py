#!/usr/bin/env python
"""One K1 launch on a 16-float (or argv[2]) column slice of the products-shaped graph -- the command ncu wraps:
  ncu --set full --import-source on -k regex:spmm_rows -c 2 -o gpurun_out/k1_narrow python tools/k1_narrow_profile.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cleora_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "products-2.4m-62m-d256"
ds = int(sys.argv[2]) if len(sys.argv) > 2 else 16
w = bench.WORKLOADS[name]
g, _ = bench.build_host_graph(w)
L = _lib.lib()
n, d = g.num_entities, w["d"]
_lib.check(L.cleora_dev_graph_prepare(g._handle()))
st = torch.cuda.current_stream().cuda_stream
xs = torch.randn(n, ds, device="cuda")
out = torch.empty(n, d, device="cuda")
dst = (C.c_void_p * 1)(out.data_ptr())
for _ in range(3):
    _lib.check(L.cleora_dev_spmm_scatter(g._handle(), 0, xs.data_ptr(), ds, dst, 1, n, d, 0, None, 1.0, 0.0, st))
torch.cuda.synchronize()
print("done")
