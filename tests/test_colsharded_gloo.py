"""Host-side logic of the column-sharded multi-GPU loop (cleora_b200/colsharded.py) on CPU: world_size 2 and 4 over
gloo.  The device kernels are replaced by an oracle-backed double and the CUDA-IPC peer buffers by files under
/dev/shm that every rank maps, so what is verified here is the layout arithmetic (column slices, row blocks, the seed
offset of the sliced init), which rank writes which rows / slices of whose buffer, and the synchronisation points --
against the single-process oracle.  The real kernels run the same choreography in tests/test_gpu_sharded.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cleora_b200 as cb
import oracle
from cleora_b200 import _lib, colsharded
from tests.helpers import KARATE_COLUMNS, KARATE_EDGES, er_lines, gram_err, procrustes_err


class ShmPeerBuffer:
    def __init__(self, tag, name, rows, cols, rank, world):
        self.paths = [f"/dev/shm/{tag}_{name}_{r}.bin" for r in range(world)]
        own = np.memmap(self.paths[rank], dtype=np.float32, mode="w+", shape=(max(rows, 1), cols))
        own[:] = 0
        dist.barrier()
        self.dests = [own if r == rank else np.memmap(self.paths[r], dtype=np.float32, mode="r+", shape=(max(rows, 1), cols))
                      for r in range(world)]
        self.tensor = torch.from_numpy(own)
        self.rank = rank

    def close(self):
        dist.barrier()
        try:
            os.unlink(self.paths[self.rank])
        except OSError:
            pass


class OracleColumnBackend:
    def __init__(self, tag, og):
        self.torch, self.tag, self.og = torch, tag, og

    def empty(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def from_numpy(self, a):
        return torch.from_numpy(np.ascontiguousarray(a))

    def peer_buffer(self, name, rows, cols, dist_, group, rank, world):
        return ShmPeerBuffer(self.tag, name, rows, cols, rank, world)

    def graph_hashes(self, graph):
        return self.og.hashes, self.og.n

    def init_slice(self, hashes, n, ds, seed, x):
        x.numpy()[:n] = oracle.init_matrix(hashes, ds, seed)

    def row_scale_full(self, graph, markov, out):
        v = self.og.values("left" if markov == 0 else "symmetric")
        out.numpy()[:self.og.n] = np.add.reduceat(np.concatenate([v, [0]]).astype(np.float32), self.og.rowptr[:-1])[:self.og.n]

    def spmm_scatter(self, graph, markov, x_slice, ds, wa, block, d, col_off, resid, alpha, rw):
        og = self.og
        xs = np.ascontiguousarray(x_slice.numpy()[:og.n])
        out = np.zeros((og.n, ds), np.float32)
        oracle.lib().orc_spmm(og.n, og.rowptr, og.col, og.values("left" if markov == 0 else "symmetric"), xs, ds, out)
        if resid is not None:
            out = np.float32(alpha) * out + np.float32(rw) * resid.numpy()[:og.n]
        for h, dst in enumerate(wa.dests):
            a, b = h * block, min(og.n, (h + 1) * block)
            if a < b:
                dst[:b - a, col_off:col_off + ds] = out[a:b]
                if isinstance(dst, np.memmap):
                    dst.flush()

    @staticmethod
    def _norm(x, norm):
        if norm == _lib.NORM_L2_RUST:
            return oracle.l2_normalize(x)
        if norm == _lib.NORM_L2_NUMPY:
            return oracle.normalize(x, "l2")
        if norm == _lib.NORM_L1_NUMPY:
            return oracle.normalize(x, "l1")
        return x

    @staticmethod
    def _slices(rows, out, xb, row_base, n):
        out.numpy()[:n] = rows
        if xb is None:
            return
        ds = rows.shape[1] // len(xb.dests)
        for h, dst in enumerate(xb.dests):
            dst[row_base:row_base + n] = rows[:, h * ds:(h + 1) * ds]
            if isinstance(dst, np.memmap):
                dst.flush()

    def normalize_slices(self, x, n, d, norm, out, xb, row_base):
        self._slices(self._norm(np.array(x.numpy()[:n]), norm), out, xb, row_base, n)

    def apply_slices(self, x, n, d, mean32, T, out, xb, row_base, norm, rowscale, t_upper=False):
        rs = np.ones((n, 1), np.float32) if rowscale is None else rowscale.numpy()[:n, None]
        q = (np.array(x.numpy()[:n]) - rs * mean32.numpy()) @ T.numpy()
        self._slices(self._norm(q, norm), out, xb, row_base, n)

    def col_sums(self, y, n, d, sums):
        sums.numpy()[:] = y.numpy()[:n].sum(axis=0, dtype=np.float64)

    def gram(self, y, n, d, mean, cov):
        b = y.numpy()[:n].astype(np.float64) - mean.numpy()
        cov.numpy()[:] = b.T @ b

    def apply(self, y, n, d, mean32, T, z):
        z.numpy()[:n] = (y.numpy()[:n] - mean32.numpy()) @ T.numpy()

    def sq_diff(self, a, b, count, f64, out):
        aa, bb = a.numpy().reshape(-1)[:count], b.numpy().reshape(-1)[:count]
        dl = aa.astype(np.float64) - bb.astype(np.float64) if f64 else (aa - bb)
        out.numpy()[0] = float(np.sum((dl * dl).astype(np.float64)))

    def transform(self, cov, d, T):
        T.numpy()[:] = oracle.whiten_transform(cov.numpy())

    def chol(self, cov, d, T, status):
        T.numpy()[:] = np.linalg.inv(np.linalg.cholesky(cov.numpy())).T.astype(np.float32)

    def chol_enabled(self, d):
        return True

    def tc_apply_ok(self, d):
        return True

    class _S:
        def wait_stream(self, other):
            pass

    def new_stream(self):
        return OracleColumnBackend._S()

    def on(self, stream):
        import contextlib
        return contextlib.nullcontext()

    def current(self):
        return OracleColumnBackend._S()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CLEORA_B200_COLSHARD="force")   # small d in these tests
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lines, columns, kw = case
        g = cb.SparseMatrix.from_iterator(lines, columns)
        og = oracle.build_graph(lines, columns)
        d = kw["feature_dim"]
        assert colsharded.eligible(d, world)
        em = colsharded.ColumnShardedEmbedder(g, d, rank, world, backend=OracleColumnBackend(f"cleora_cs_{port}", og))
        whiten = kw.get("whiten", True)
        rust = not whiten
        norm = _lib.NORM_L2_RUST if rust else _lib.NORM_L2_NUMPY
        rw, thr = kw.get("residual_weight", 0.0), kw.get("convergence_threshold", 0.0)
        markov = _lib.MARKOV[kw.get("propagation", "left")]
        if em.pipeline_eligible(kw["num_iterations"], norm, rw, thr, whiten):
            em.run_pipelined(kw["num_iterations"], markov, 0, None)
        else:
            em.run(kw["num_iterations"], markov, norm, 0, None, rw, thr, whiten, rust_semantics=rust)
        out = em.result()
        em.close()
        if rank == 0:
            ret.put(out)
    finally:
        dist.destroy_process_group()


def run_cols(world, case):
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, ret)) for r in range(world)]
    for p in procs:
        p.start()
    out = None
    for _ in range(240):
        if not ret.empty():
            out = ret.get()
            break
        if any(p.exitcode not in (None, 0) for p in procs):
            break
        procs[0].join(0.5)
    for p in procs:
        p.join(60)
    codes = [p.exitcode for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert codes == [0] * world and out is not None, f"worker exit codes {codes}"
    return out


ER = (er_lines(601, 4000, 3), "complex::reflexive::node")          # 601 rows: the last row block is shorter


def test_eligibility(monkeypatch):
    assert colsharded.eligible(256, 8) and colsharded.eligible(128, 4) and colsharded.eligible(256, 2)
    assert not colsharded.eligible(128, 8)                  # 16-float slices: row-sharded is faster (HBM granularity)
    monkeypatch.setenv("CLEORA_B200_COLSHARD", "force")
    assert colsharded.eligible(128, 8) and colsharded.eligible(32, 4)
    monkeypatch.delenv("CLEORA_B200_COLSHARD")
    assert not colsharded.eligible(256, 1) and not colsharded.eligible(48, 4) and not colsharded.eligible(100, 2)
    assert not colsharded.eligible(64, 3)


@pytest.mark.parametrize("world", [2, 4])
def test_column_sharded_fast_path_is_bit_identical_to_single_process(world):
    kw = dict(feature_dim=32, num_iterations=6, whiten=False, residual_weight=0.25)
    out = run_cols(world, (ER[0], ER[1], kw))
    ref = oracle.embed(oracle.build_graph(ER[0], ER[1]), **kw)
    np.testing.assert_array_equal(out, ref)


def test_column_sharded_whitened_loop_matches_oracle_karate():
    kw = dict(feature_dim=16, num_iterations=5)                              # d % 32 != 0: reference stage order
    out = run_cols(2, (KARATE_EDGES, KARATE_COLUMNS, kw))
    ref = oracle.embed(oracle.build_graph(KARATE_EDGES, KARATE_COLUMNS), **kw)
    sign = np.sign(np.sum(out * ref, axis=0))
    assert np.max(np.abs(out * sign - ref)) <= 1e-4 * np.max(np.abs(ref))


@pytest.mark.parametrize("world", [2, 4])
def test_column_sharded_reference_order_with_residual(world):
    kw = dict(feature_dim=32, num_iterations=5, residual_weight=0.1)         # residual: not pipelined
    out = run_cols(world, (ER[0], ER[1], kw))
    ref = oracle.embed(oracle.build_graph(ER[0], ER[1]), **kw)
    assert gram_err(out, ref) <= 1e-4
    assert procrustes_err(out, ref) <= 1e-3


@pytest.mark.parametrize("world", [2, 4])
def test_column_sharded_pipelined_choreography_matches_oracle(world):
    kw = dict(feature_dim=32, num_iterations=5)
    out = run_cols(world, (ER[0], ER[1], kw))
    ref = oracle.embed(oracle.build_graph(ER[0], ER[1]), **kw)
    assert gram_err(out, ref) <= 1e-4
    assert procrustes_err(out, ref) <= 1e-3


def test_column_sharded_convergence_and_symmetric():
    kw = dict(feature_dim=16, num_iterations=30, whiten=False, convergence_threshold=0.02, propagation="symmetric")
    out = run_cols(2, (ER[0], ER[1], kw))
    og = oracle.build_graph(ER[0], ER[1])
    ref, it = oracle.embed_fast_convergence(og, 16, 30, "symmetric", 0, 0.0, 0.02)
    assert it < 30
    np.testing.assert_array_equal(out, ref)
