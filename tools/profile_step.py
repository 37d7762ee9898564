#!/usr/bin/env python
"""One (or a few) iterations of the default embed loop on a bench workload -- the command ncu wraps.
  ncu --set full -k regex:"spmm_rows|gram_f64|whiten_apply" -c 3 -o gpurun_out/prof python tools/profile_step.py
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import cleora_b200 as cb  # noqa: E402
from cleora_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="er-1m-20m-d256")
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--whiten", type=int, default=1)
a = ap.parse_args()
w = bench.WORKLOADS[a.workload]
u, v = bench.gen_pairs(w)
g = cb.SparseMatrix.from_edge_arrays(u, v)
t = np.zeros(8)
out, _ = g.embed_device(w["d"], a.iters, "left", _lib.NORM_L2_NUMPY if a.whiten else _lib.NORM_L2_RUST, 0, None, 0.0,
                        0.0, bool(a.whiten), timings=t)
print("phases ms (h2d init spmm stats eigh apply rmse d2h):", np.round(t, 3), "launches",
      _lib.lib().cleora_kernel_launch_count())
