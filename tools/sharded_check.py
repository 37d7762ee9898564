#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun on N GPUs of one box:
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py
Row-sharded embed vs the single-GPU path of the same library (rank 0) and vs the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cleora_b200 as cb  # noqa: E402
from cleora_b200 import sharded  # noqa: E402
from tests.helpers import er_lines  # noqa: E402
from tests.helpers import gram_err, procrustes_err  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lines = er_lines(30000, 400000, 5)
g = cb.SparseMatrix.from_iterator(lines, "complex::reflexive::node")
ok = True
for kw in (dict(feature_dim=64, num_iterations=10, whiten=False),
           dict(feature_dim=128, num_iterations=6, whiten=False, residual_weight=0.2, propagation="symmetric"),
           dict(feature_dim=64, num_iterations=5, whiten=True),
           dict(feature_dim=256, num_iterations=3, whiten=True)):
    out = sharded.embed_sharded(g, **kw)
    if rank == 0:
        ref = cb.embed(g, **kw)                       # single-GPU path, same library
        if not kw["whiten"]:
            same = np.array_equal(out, ref)
            print(f"[{world} GPUs] {kw}: bit-identical to 1 GPU: {same}", flush=True)
            ok &= same
        else:
            ge, pe = gram_err(out, ref), procrustes_err(out, ref)
            print(f"[{world} GPUs] {kw}: gram err {ge:.2e}, procrustes err {pe:.2e}", flush=True)
            ok &= ge < 1e-5 and pe < 1e-4
dist.barrier()
if rank == 0:
    print("SHARDED_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
