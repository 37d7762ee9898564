"""Row-sharded loop (cleora_b200/sharded.py) with the REAL device backend, several ranks, on the GPU box (`-m gpu`).

Two launch modes, same worker:
  * "one-gpu": WORLD ranks share cuda:0 and talk over gloo (CUDA tensors).  This always runs -- also on the driver's
    single-GPU box -- and exercises everything but NCCL itself: nnz-balanced shards with remapped columns, the padded
    gathered layout, the CUDA-IPC peer matrix and the fused gather of K1's epilogue (peer stores into another
    process's buffer), the three-stream pipelined choreography, the all-reduced statistics, the Cholesky-whitened inner
    iterations computed redundantly per rank, the final PCA transform + broadcast.
  * "nccl": one rank per visible GPU (2..8), NCCL over NVLink; skipped when fewer than two GPUs are visible.
embed_sharded() picks the column-sharded loop (cleora_b200/colsharded.py: SpMM on column slices, both transposes fused
into kernel epilogues as peer stores) when d splits evenly into supported slices -- world 2, 4, 8 for most cases below --
and the row-sharded loop with its all-gather otherwise (world 3, d = 48), so both implementations are exercised.
Bars: whiten=False is bit-identical to the single-GPU path for every rank count (accumulation order inside a row does
not depend on the sharding); the whitened loop agrees with the single-GPU path and with the CPU oracle in
Procrustes (<= 1e-4) and Gram (<= 1e-5) terms."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    dict(feature_dim=64, num_iterations=10, whiten=False),
    dict(feature_dim=128, num_iterations=6, whiten=False, residual_weight=0.2, propagation="symmetric"),
    dict(feature_dim=64, num_iterations=8, whiten=True),                     # pipelined, Cholesky inner iterations
    dict(feature_dim=256, num_iterations=6, whiten=True),
    dict(feature_dim=48, num_iterations=5, whiten=True),                     # reference stage order (d % 32 != 0)
    dict(feature_dim=32, num_iterations=12, whiten=True, convergence_threshold=1e-9),   # rmse path: PCA every iteration
]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _graph_lines():
    from tests.helpers import er_lines
    return er_lines(30000, 400000, 5), "complex::reflexive::node"


def _worker(rank, world, port, backend, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cleora_b200 as cb
        from cleora_b200 import sharded
        lines, columns = _graph_lines()
        g = cb.SparseMatrix.from_iterator(lines, columns)
        outs = [sharded.embed_sharded(g, **kw) for kw in CASES]
        if rank == 0:
            ret.put(outs)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run(world, backend):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, ret)) for r in range(world)]
    for p in procs:
        p.start()
    outs = None
    for _ in range(600):
        if not ret.empty():
            outs = ret.get()
            break
        if any(p.exitcode not in (None, 0) for p in procs):
            break
        procs[0].join(1.0)
    for p in procs:
        p.join(120)
    codes = [p.exitcode for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert codes == [0] * world and outs is not None, f"worker exit codes {codes}"
    return outs


def _check(outs):
    import cleora_b200 as cb
    import oracle
    from tests.helpers import gram_err, procrustes_err
    lines, columns = _graph_lines()
    g = cb.SparseMatrix.from_iterator(lines, columns)
    o = oracle.build_graph(lines, columns)
    for kw, out in zip(CASES, outs):
        single = cb.embed(g, **kw)
        if not kw["whiten"]:
            np.testing.assert_array_equal(out, single, err_msg=str(kw))          # bit-identical to one GPU
        else:
            ref = oracle.embed(o, **kw)
            assert out.shape == ref.shape == single.shape, kw
            for other, name in ((single, "single GPU"), (ref, "oracle")):
                assert procrustes_err(out, other) <= 1e-4, (kw, name)
                assert gram_err(out, other) <= 1e-5, (kw, name)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_sharded_ranks_sharing_one_gpu(world):
    _check(_run(world, "gloo"))


def test_sharded_nccl_all_visible_gpus():
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus N)")
    _check(_run(n, "nccl"))
