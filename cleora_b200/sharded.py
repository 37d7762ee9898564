"""Row-sharded multi-GPU embed loop: one process per GPU, ``torch.distributed`` (NCCL over NVLink) for the
plumbing, the library's device-level kernels (``cleora_dev_*``) for every numerical step.

Partitioning: contiguous row blocks balanced by nnz.  Rank g owns rows [r_g, r_{g+1}) of the CSR,
of the iterate and of the output.  The two couplings of the path are its only collectives:
  1. K1 reads rows of X at arbitrary columns  -> all-gather of the owned block before each SpMM;
  2. whitening needs the global mean and d x d covariance -> all-reduce of d + d*d f64 partials; ``eigh`` once on
     rank 0, T (d x d f32) broadcast.
The gathered matrix uses a padded layout -- block g starts at row g*B with B = max block size -- so equal-sized
NCCL all-gathers work for unequal blocks; shard column indices are remapped to that layout once, at shard
creation (padding rows are never referenced).  Accumulation order inside a row is unchanged, so every rank count
produces the same SpMM bits as one GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Optional

import numpy as np

from . import _lib
from ._lib import check
from .pycleora import SparseMatrix


# ------------------------------------------------------------------------------------------------ partitioning
def partition_rows_by_nnz(rowptr: np.ndarray, world: int) -> np.ndarray:
    """Boundaries r_0=0 <= r_1 <= ... <= r_world=n with ~equal nnz per block (each block non-empty when n>=world)."""
    n = rowptr.shape[0] - 1
    nnz = int(rowptr[-1])
    # weight = nnz + rows, so that edge-less stretches still get split
    w = rowptr.astype(np.float64) + np.arange(n + 1, dtype=np.float64)
    targets = (nnz + n) * np.arange(1, world, dtype=np.float64) / world
    cuts = np.searchsorted(w, targets, side="left")
    bounds = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    for g in range(1, world + 1):                         # strictly increasing where possible
        bounds[g] = max(bounds[g], min(bounds[g - 1] + 1, n))
    bounds[world] = n
    for g in range(world - 1, 0, -1):
        bounds[g] = min(bounds[g], bounds[g + 1])
    return bounds


def padded_index(cols: np.ndarray, bounds: np.ndarray, block: int) -> np.ndarray:
    """Global row/column index -> index in the padded gathered layout (owner*block + offset in owner)."""
    owner = np.searchsorted(bounds, cols, side="right") - 1
    return (owner.astype(np.int64) * block + (cols.astype(np.int64) - bounds[owner])).astype(np.uint32)


class Shard:
    """This rank's rows of the operator, with columns in the padded layout."""

    def __init__(self, rowptr, col, left, sym, hashes, rank: int, world: int):
        n = rowptr.shape[0] - 1
        self.n, self.rank, self.world = n, rank, world
        self.bounds = partition_rows_by_nnz(np.asarray(rowptr), world)
        self.block = int(np.max(np.diff(self.bounds))) if n else 0
        self.r0, self.r1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.n_local = self.r1 - self.r0
        e0, e1 = int(rowptr[self.r0]), int(rowptr[self.r1])
        self.nnz_local, self.nnz = e1 - e0, int(rowptr[-1])
        self.n_pad = self.block * world
        local_rowptr = np.asarray(rowptr[self.r0:self.r1 + 1], dtype=np.int64) - e0
        local_col = padded_index(np.asarray(col[e0:e1]), self.bounds, self.block)
        self.graph = SparseMatrix.from_csr(local_rowptr, local_col, np.asarray(left[e0:e1]),
                                           None if sym is None else np.asarray(sym[e0:e1]), None, None,
                                           n_cols=max(self.n_pad, 1), row_offset=self.rank * self.block)
        # entity hashes of ALL rows in padded order (init is computed for the whole gathered matrix)
        hp = np.zeros(self.n_pad, np.uint64)
        if n:
            hp[padded_index(np.arange(n, dtype=np.int64), self.bounds, self.block)] = np.asarray(hashes, np.uint64)
        self.hash_padded = hp

    @classmethod
    def from_device_pairs(cls, u, v, rank: int, world: int, want_sym: bool = False, column_name: str = "node") -> "Shard":
        """This rank's row block built on the GPU straight from the (device-resident) pair arrays -- every rank passes
        the same pairs; nothing of the CSR touches the host (cleora_dev_graph_from_pairs)."""
        self = cls.__new__(cls)
        g = SparseMatrix.from_edge_arrays_device(u, v, column_name, rank, world, want_sym)
        L = _lib.lib()
        self.graph, self.rank, self.world = g, rank, world
        self.bounds = np.asarray(g.shard_bounds, np.int64)
        self.n = int(L.cleora_graph_num_entities_global(g._handle()))
        self.block = int(np.max(np.diff(self.bounds))) if self.n else 0
        self.r0, self.r1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.n_local = self.r1 - self.r0
        self.nnz_local = g.num_edges
        self.nnz = None                                   # global nnz: sum over ranks (bench all-reduces it)
        self.n_pad = self.block * world
        hp, nh = C.c_void_p(), C.c_int64()
        check(L.cleora_dev_graph_hashes(g._handle(), C.byref(hp), C.byref(nh)))
        self.hash_padded = None
        self.hash_ptr = hp.value                          # device pointer, n_pad entries (library-owned)
        assert world == 1 or int(nh.value) == self.n_pad
        return self

    @classmethod
    def from_matrix(cls, g: SparseMatrix, rank: int, world: int) -> "Shard":
        rowptr, col, left, sym = g._csr()
        return cls(rowptr, col, left, sym, g.entity_hashes(), rank, world)

    def unpad(self, x_pad: np.ndarray) -> np.ndarray:
        out = np.empty((self.n, x_pad.shape[1]), x_pad.dtype)
        for g in range(self.world):
            a, b = int(self.bounds[g]), int(self.bounds[g + 1])
            out[a:b] = x_pad[g * self.block:g * self.block + (b - a)]
        return out

    def pad(self, x: np.ndarray) -> np.ndarray:
        out = np.zeros((self.n_pad, x.shape[1]), x.dtype)
        for g in range(self.world):
            a, b = int(self.bounds[g]), int(self.bounds[g + 1])
            out[g * self.block:g * self.block + (b - a)] = x[a:b]
        return out


# ------------------------------------------------------------------------------------------------ device backend
class CudaBackend:
    """Device-level ABI calls on torch-owned memory (torch = allocator + streams + NCCL only)."""

    def __init__(self, device: int):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(device)
        self.L = _lib.lib()
        check(self.L.cleora_set_device(device))

    def stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def empty(self, shape, dtype):
        return self.torch.empty(shape, dtype=dtype, device=self.device)

    def from_numpy(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def init(self, hash_pad, n_pad, d, seed, x):
        check(self.L.cleora_dev_init(hash_pad.data_ptr(), n_pad, d, seed, x.data_ptr(), self.stream()))

    def spmm(self, shard, markov, x_full, d, y, resid, alpha, rw, norm):
        check(self.L.cleora_dev_spmm(shard.graph._handle(), markov, x_full.data_ptr(), d, y.data_ptr(),
                                     None if resid is None else resid.data_ptr(), alpha, rw, norm, self.stream()))

    def col_sums(self, y, n, d, sums):
        check(self.L.cleora_dev_col_sums(y.data_ptr(), n, d, sums.data_ptr(), 0, self.stream()))

    def gram(self, y, n, d, mean, cov):
        check(self.L.cleora_dev_centered_gram(y.data_ptr(), n, d, mean.data_ptr(), cov.data_ptr(), self.stream()))

    def apply(self, y, n, d, mean32, T, z):
        check(self.L.cleora_dev_whiten_apply(y.data_ptr(), n, d, mean32.data_ptr(), T.data_ptr(), d, z.data_ptr(),
                                             self.stream()))

    def sq_diff(self, a, b, count, f64, out):
        check(self.L.cleora_dev_sq_diff_sum(a.data_ptr(), b.data_ptr(), count, 1 if f64 else 0, out.data_ptr(),
                                            self.stream()))

    def transform(self, cov, d: int, T) -> None:
        """cov (device f64, scaled) -> T (device f32) on the current stream (cuSOLVER unless a host eigh is set)."""
        check(self.L.cleora_dev_whiten_transform(cov.data_ptr(), d, d, T.data_ptr(), self.stream()))

    def chol(self, cov, d: int, T, status) -> None:
        """cov (device f64, scaled) -> T = L^-T (device f32) on the current stream; status (int32[1]) is raised when cov
        is not safely positive definite (chol_whiten.cu)."""
        check(self.L.cleora_dev_chol_whiten(cov.data_ptr(), d, T.data_ptr(), status.data_ptr(), self.stream()))

    def chol_enabled(self, d: int) -> bool:
        return d <= 512 and self.L.cleora_get_option(b"chol_whiten") == 1

    def fusable(self, d: int) -> bool:
        return bool(self.L.cleora_whiten_apply_fusable(d, d))

    def apply_ex(self, x, n, d, mean32, T, out, norm, rowscale, t_upper=False):
        check(self.L.cleora_dev_whiten_apply_ex(x.data_ptr(), n, d, mean32.data_ptr(), T.data_ptr(), d, out.data_ptr(),
                                                norm, None if rowscale is None else rowscale.data_ptr(),
                                                1 if t_upper else 0, self.stream()))

    def row_scale(self, shard, markov, out):
        check(self.L.cleora_dev_row_scale(shard.graph._handle(), markov, out.data_ptr(), self.stream()))

    # ---- fused all-gather: kernels store their rows into every rank's copy of the gathered matrix (CUDA IPC)
    def _ptr_array(self, ptrs):
        arr = (C.c_void_p * max(len(ptrs), 1))(*ptrs)
        return arr

    def spmm_push(self, shard, markov, x_full, d, out_ptr, extra_ptrs, resid, alpha, rw, norm):
        check(self.L.cleora_dev_spmm_push(shard.graph._handle(), markov, x_full.data_ptr(), d, out_ptr,
                                          self._ptr_array(extra_ptrs), len(extra_ptrs),
                                          None if resid is None else resid.data_ptr(), alpha, rw, norm, self.stream()))

    def apply_push(self, x, n, d, mean32, T, out_ptr, extra_ptrs, norm, rowscale):
        check(self.L.cleora_dev_whiten_apply_push(x.data_ptr(), n, d, mean32.data_ptr(), T.data_ptr(), d, out_ptr,
                                                  self._ptr_array(extra_ptrs), len(extra_ptrs), norm,
                                                  None if rowscale is None else rowscale.data_ptr(), self.stream()))

    def peer_matrix(self, rows, d, dist, group, rank, world):
        """Two [rows, d] f32 buffers on this GPU whose addresses are mapped into every other rank (ping-pong)."""
        return PeerMatrix(self, rows, d, dist, group, rank, world)

    # streams: eigensolve / gather run beside the main stream in the pipelined loop
    def new_stream(self):
        # highest priority: the one-CTA Cholesky kernel (or cuSOLVER's chain of small kernels) must get an SM slot ahead
        # of the SpMM's ~100k CTAs that run beside it (default priority: 9 ms instead of 1 ms at C3 on 2 GPUs, measured)
        return self.torch.cuda.Stream(device=self.device, priority=-1)

    def on(self, stream):
        return self.torch.cuda.stream(stream)

    def current(self):
        return self.torch.cuda.current_stream()

    def sync(self):
        self.torch.cuda.synchronize()


class _DevPtr:
    """A raw device address with the one method the backend needs."""

    def __init__(self, address: int):
        self._a = int(address)

    def data_ptr(self) -> int:
        return self._a


class _CudaArray:
    """__cuda_array_interface__ shim so torch can view library-allocated (IPC-exportable) device memory."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerMatrix:
    """The gathered iterate in peer-writable memory: buf[b] is this rank's copy b (b = 0, 1), peer[b][p] the address
    of rank p's copy b as mapped into this process (cudaIpcOpenMemHandle with lazy peer access)."""

    def __init__(self, be, rows, d, dist, group, rank, world):
        self.be, self.rank, self.world = be, rank, world
        L = be.L
        nbytes = max(rows * d * 4, 4)
        self.ptr, handles = [], []
        for _ in range(2):
            p = C.c_void_p()
            check(L.cleora_dev_malloc(nbytes, C.byref(p)))
            h = C.create_string_buffer(64)
            check(L.cleora_ipc_get_handle(p, h))
            self.ptr.append(p.value)
            handles.append(h.raw)
        everyone = [None] * world
        dist.all_gather_object(everyone, handles, group=group)
        self.peer, self._opened = [[None] * world for _ in range(2)], []
        for b in range(2):
            for p in range(world):
                if p == rank:
                    self.peer[b][p] = self.ptr[b]
                else:
                    q = C.c_void_p()
                    check(L.cleora_ipc_open(everyone[p][b], C.byref(q)))
                    self.peer[b][p] = q.value
                    self._opened.append(q.value)
        self.buf = [be.torch.as_tensor(_CudaArray(self.ptr[b], (rows, d)), device=be.device) for b in range(2)]

    def block_ptrs(self, b, block_row0, d):
        """(own address, [peer addresses]) of the row block starting at block_row0 in copy b."""
        off = block_row0 * d * 4
        return self.peer[b][self.rank] + off, [self.peer[b][p] + off for p in range(self.world) if p != self.rank]

    def close(self):
        L = self.be.L
        for q in self._opened:
            L.cleora_ipc_close(q)
        self._opened = []
        for p in self.ptr:
            L.cleora_dev_free(p)
        self.ptr = []


# ------------------------------------------------------------------------------------------------ the loop
class ShardedEmbedder:
    """Device-resident, row-sharded version of embed() (pycleora/__init__.py:51-127)."""

    def __init__(self, shard: Shard, d: int, backend=None, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.shard, self.d, self.group = shard, d, group
        self.be = backend if backend is not None else CudaBackend(torch.cuda.current_device())
        be, s = self.be, shard
        self.pm = None
        self.cur = 0
        if (hasattr(be, "peer_matrix") and 1 < s.world <= 8 and s.n_pad > 0 and d in (8, 16, 32, 64, 96, 128, 192, 256, 384, 512, 1024)
                and os.environ.get("CLEORA_B200_P2P", "1") != "0"):
            try:
                self.pm = be.peer_matrix(s.n_pad, d, dist, group, s.rank, s.world)
            except (RuntimeError, ValueError):
                self.pm = None
        ok = torch.tensor([1 if self.pm is not None else 0])
        if s.world > 1 and hasattr(be, "peer_matrix"):
            okd = ok.to(be.device)
            dist.all_reduce(okd, op=dist.ReduceOp.MIN, group=group)          # all ranks or none
            if int(okd.item()) == 0 and self.pm is not None:
                self.pm.close()
                self.pm = None
        # gathered iterate, padded layout (copy `cur` of the peer-writable pair when the fused gather is available)
        self.x_full = self.pm.buf[0] if self.pm is not None else be.empty((max(s.n_pad, 1), d), torch.float32)
        self.x_prev = None
        self.y = be.empty((max(s.block, 1), d), torch.float32)           # own block, post SpMM+norm
        self.z = be.empty((max(s.block, 1), d), torch.float32)           # own block, post whitening
        self.sums = be.empty((d,), torch.float64)
        self.cov = be.empty((d, d), torch.float64)
        self.mean32 = be.empty((d,), torch.float32)
        self.T = be.empty((d, d), torch.float32)
        self.scalar = be.empty((1,), torch.float64)
        self.status = be.empty((1,), torch.int32)               # raised by the Cholesky whitening kernel
        self.status.zero_()
        if getattr(s, "hash_ptr", None):
            self.hash_pad = _DevPtr(s.hash_ptr)                                # device-built shard: hashes already in HBM
        else:
            self.hash_pad = be.from_numpy(s.hash_padded.view(np.int64)) if s.n_pad else be.empty((1,), torch.int64)
        self.phase_ms = {}
        if s.n_pad:
            self.x_full.zero_()
        self.y.zero_()
        self.z.zero_()

    def _own(self, t):
        s = self.shard
        return t[s.rank * s.block:(s.rank + 1) * s.block]

    def _gather(self, block):
        # equal-sized blocks: one NCCL all-gather into the padded matrix
        self.dist.all_gather_into_tensor(self.x_full, block, group=self.group)

    def _push_targets(self):
        """Addresses of this rank's block in the NEXT copy of the gathered matrix, here and on every peer."""
        s = self.shard
        return self.pm.block_ptrs(1 - self.cur, s.rank * s.block, self.d)

    def _flip(self, barrier_group=None):
        """All ranks have pushed their blocks: make the next copy current (tiny all-reduce = node-wide barrier)."""
        if not hasattr(self, "_flag"):
            self._flag = self.be.empty((1,), self.torch.float32)
            self._flag.zero_()
        self.dist.all_reduce(self._flag, group=barrier_group if barrier_group is not None else self.group)
        self.cur = 1 - self.cur
        self.x_full = self.pm.buf[self.cur]

    def run(self, iters: int, markov: int = 0, norm: int = _lib.NORM_L2_NUMPY, seed: int = 0,
            x0: Optional[np.ndarray] = None, residual_weight: float = 0.0, convergence_threshold: float = 0.0,
            whiten: bool = True, rust_semantics: bool = False, timers=None, _allow_chol: bool = True) -> int:
        """Leaves the final iterate in self.x_full (padded layout) on every rank; returns iterations done."""
        torch, dist, be, s, d = self.torch, self.dist, self.be, self.shard, self.d
        n = s.n
        if x0 is not None:
            self.x_full.copy_(be.from_numpy(s.pad(np.ascontiguousarray(x0, np.float32))))
        elif s.n_pad:
            be.init(self.hash_pad, s.n_pad, d, seed, self.x_full)
        if rust_semantics:                                   # src/embedding.rs:116
            use_res = 0.0 < residual_weight < 1.0
            alpha, rw = float(np.float32(1.0) - np.float32(residual_weight)), float(np.float32(residual_weight))
        else:                                                # pycleora/__init__.py:114
            use_res = residual_weight > 0
            alpha, rw = float(np.float32(1.0 - residual_weight)), float(np.float32(residual_weight))
        conv = convergence_threshold > 0
        do_whiten = whiten and n > 1
        if conv and self.x_prev is None:
            self.x_prev = be.empty(tuple(self.x_full.shape), torch.float32)
        # iterates that never leave the loop are whitened with the Cholesky factor, computed redundantly (and
        # identically: the all-reduced covariance is bit-identical everywhere) on every rank -- no eigensolve, no
        # broadcast; the last iterate gets the reference's PCA transform (see chol_whiten.cu for the argument)
        inner_chol = (_allow_chol and do_whiten and not conv and iters >= 2 and hasattr(be, "chol") and be.chol_enabled(d)
                      and norm in (_lib.NORM_L2_NUMPY, _lib.NORM_NONE))
        if inner_chol:
            self.status.zero_()
        done = 0
        for it in range(iters):
            if conv:
                self.x_prev.copy_(self.x_full)
            t = timers.start("spmm") if timers else None
            pushed = self.pm is not None and not do_whiten
            if pushed:      # K1 stores its rows straight into every rank's next copy: the all-gather is the epilogue
                own_ptr, extra = self._push_targets()
                be.spmm_push(s, markov, self.x_full, d, own_ptr, extra, self._own(self.x_full) if use_res else None,
                             alpha, rw, norm)
            else:
                be.spmm(s, markov, self.x_full, d, self.y, self._own(self.x_full) if use_res else None, alpha, rw, norm)
            if timers:
                timers.stop(t)
            fresh = self.y
            if do_whiten:
                t = timers.start("stats") if timers else None
                be.col_sums(self.y, s.n_local, d, self.sums)
                dist.all_reduce(self.sums, group=self.group)
                self.sums.div_(float(n))                                        # mean (f64)
                be.gram(self.y, s.n_local, d, self.sums, self.cov)
                dist.all_reduce(self.cov, group=self.group)
                self.cov.mul_(1.0 / float(n - 1))
                self.mean32.copy_(self.sums)                                    # astype(float32)
                if timers:
                    timers.stop(t)
                t = timers.start("eigh") if timers else None
                if inner_chol and it + 1 < iters:
                    be.chol(self.cov, d, self.T, self.status)
                else:
                    if s.rank == 0:
                        be.transform(self.cov, d, self.T)
                    dist.broadcast(self.T, src=0, group=self.group)   # one eigensolve, identical T everywhere
                if timers:
                    timers.stop(t)
                t = timers.start("apply") if timers else None
                be.apply(self.y, s.n_local, d, self.mean32, self.T, self.z)
                if timers:
                    timers.stop(t)
                fresh = self.z
            t = timers.start("gather") if timers else None
            if pushed:
                self._flip()
            else:
                self._gather(fresh)
            if timers:
                timers.stop(t)
            done = it + 1
            if conv and it > 0:
                be.sq_diff(self._own(self.x_full), self._own(self.x_prev), s.n_local * d, not rust_semantics, self.scalar)
                dist.all_reduce(self.scalar, group=self.group)
                tot = float(self.scalar.item())
                if rust_semantics:
                    rmse = float(np.sqrt(np.float32(tot) / np.float32(n * d)))
                else:
                    rmse = float(np.sqrt(tot / (n * d)))
                if rmse < convergence_threshold:
                    break
        if inner_chol and not self._chol_status_ok():
            return self.run(iters, markov, norm, seed, x0, residual_weight, convergence_threshold, whiten,
                            rust_semantics, timers, _allow_chol=False)
        return done

    def _chol_status_ok(self) -> bool:
        """No rank's Cholesky step flagged its covariance (identical inputs, so in practice all or none)."""
        flag = self.status.clone()
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MAX, group=self.group)
        return int(flag.item()) == 0

    # ------------------------------------------------------------------------------------------ pipelined variant
    def pipeline_eligible(self, iters, norm, residual_weight, convergence_threshold, whiten) -> bool:
        return (bool(whiten) and self.shard.n > 1 and iters >= 2 and norm == _lib.NORM_L2_NUMPY
                and residual_weight == 0 and convergence_threshold <= 0 and hasattr(self.be, "apply_ex")
                and self.be.fusable(self.d) and os.environ.get("CLEORA_B200_PIPELINE", "1") != "0")

    def _stats(self, y, timers=None):
        """mean (self.sums, f64), mean32, cov (scaled) of the row-distributed matrix whose local block is y."""
        be, s, d, n, dist = self.be, self.shard, self.d, self.shard.n, self.dist
        t = timers.start("stats") if timers else None
        be.col_sums(y, s.n_local, d, self.sums)
        dist.all_reduce(self.sums, group=self.group)
        self.sums.div_(float(n))
        be.gram(y, s.n_local, d, self.sums, self.cov)
        dist.all_reduce(self.cov, group=self.group)
        self.cov.mul_(1.0 / float(n - 1))
        self.mean32.copy_(self.sums)
        if timers:
            timers.stop(t)

    def run_pipelined(self, iters: int, markov: int = 0, seed: int = 0, x0: Optional[np.ndarray] = None,
                      timers=None, _allow_chol: bool = True) -> int:
        """Same mathematics as run() for the default configuration, with the eigensolve (rank 0, side stream) hidden
        behind the local SpMM via A (Y - 1 mu^T) T = (A Y - (A 1) mu^T) T, and the all-gather of the next iterate
        (own communicator, own stream) hidden behind the covariance pass.  x_full holds the gathered NORMALISED
        iterate Y between iterations; the final iterate X_T = (Y - 1 mu^T) T is gathered at the end."""
        torch, dist, be, s, d = self.torch, self.dist, self.be, self.shard, self.d
        if not hasattr(self, "_pl"):
            side, comm = be.new_stream(), be.new_stream()
            self._pl = dict(side=side, comm=comm, g_bcast=dist.new_group(), g_gather=dist.new_group(),
                            w=be.empty((max(s.block, 1), d), torch.float32),
                            rowscale=be.empty((max(s.block, 1),), torch.float32))
            self._pl["w"].zero_()
            self._pl["markov"] = None
        pl = self._pl
        side, comm, w, rowscale = pl["side"], pl["comm"], pl["w"], pl["rowscale"]
        inner_chol = _allow_chol and hasattr(be, "chol") and be.chol_enabled(d)
        if inner_chol:
            self.status.zero_()
        if pl["markov"] != markov:
            be.row_scale(s, markov, rowscale)
            pl["markov"] = markov
        main = be.current()
        if x0 is not None:
            self.x_full.copy_(be.from_numpy(s.pad(np.ascontiguousarray(x0, np.float32))))
        elif s.n_pad:
            be.init(self.hash_pad, s.n_pad, d, seed, self.x_full)
        # iteration 0: Y = rownorm(A X0); gather Y beside the stats
        # Fused gather from the tensor-core GEMM's epilogue exists (apply_push) but is off by default: that epilogue
        # writes one row per thread (16-byte pieces), which is fine for local HBM and poor over NVLink (measured: 3.5 ms
        # vs 0.5 ms per GEMM at 2 GPUs), and here the NCCL all-gather is hidden behind the covariance pass anyway.
        push = (self.pm is not None and hasattr(be, "apply_push")
                and os.environ.get("CLEORA_B200_P2P_APPLY", "0") == "1")
        t = timers.start("spmm") if timers else None
        be.spmm(s, markov, self.x_full, d, self.y, None, 1.0, 0.0, _lib.NORM_L2_NUMPY)
        if timers:
            timers.stop(t)
        y, y2 = self.y, self.z
        comm.wait_stream(main)
        with be.on(comm):
            dist.all_gather_into_tensor(self.x_full, y, group=pl["g_gather"])
        self._stats(y, timers)
        for it in range(1, iters):
            side.wait_stream(main)                                   # cov of this iterate is ready
            with be.on(side):
                t = timers.start("eigh") if timers else None
                if inner_chol:
                    be.chol(self.cov, d, self.T, self.status)        # every rank, identical input -> identical T
                else:
                    if s.rank == 0:
                        be.transform(self.cov, d, self.T)
                    dist.broadcast(self.T, src=0, group=pl["g_bcast"])
                if timers:
                    timers.stop(t)
            main.wait_stream(comm)                                   # gathered Y is complete
            t = timers.start("spmm") if timers else None
            be.spmm(s, markov, self.x_full, d, w, None, 1.0, 0.0, _lib.NORM_NONE)      # W = A Y (own rows)
            if timers:
                timers.stop(t)
            main.wait_stream(side)                                   # T has arrived
            t = timers.start("apply") if timers else None
            if push:
                # the GEMM's epilogue stores the normalised rows into y2 AND into every peer's next gathered copy;
                # y2 itself is the own block of this rank's next copy, so the stats below read it in place
                own_ptr, extra = self._push_targets()
                y2 = self._own(self.pm.buf[1 - self.cur])
                be.apply_push(w, s.n_local, d, self.mean32, self.T, own_ptr, extra, _lib.NORM_L2_NUMPY, rowscale)
            else:
                be.apply_ex(w, s.n_local, d, self.mean32, self.T, y2, _lib.NORM_L2_NUMPY, rowscale, inner_chol)
            if timers:
                timers.stop(t)
            comm.wait_stream(main)
            with be.on(comm):
                t = timers.start("gather") if timers else None
                if push:
                    self._flip(pl["g_gather"])                                         # node-wide barrier, beside the stats
                else:
                    dist.all_gather_into_tensor(self.x_full, y2, group=pl["g_gather"])  # next Y, beside the stats
                if timers:
                    timers.stop(t)
            self._stats(y2, timers)
            if not push:
                y, y2 = y2, y
            else:
                y = y2
        # final: X_T = (Y - 1 mu^T) T, gathered
        if s.rank == 0:
            be.transform(self.cov, d, self.T)
        dist.broadcast(self.T, src=0, group=self.group)
        main.wait_stream(comm)
        out_final = self.z if push else y2             # (with the fused gather y aliases the gathered matrix)
        be.apply(y, s.n_local, d, self.mean32, self.T, out_final)
        self._gather(out_final)
        if inner_chol and not self._chol_status_ok():               # a covariance was not safely SPD: eigensolver throughout
            return self.run_pipelined(iters, markov, seed, x0, timers, _allow_chol=False)
        return iters

    def result(self) -> np.ndarray:
        return self.shard.unpad(self.x_full.cpu().numpy())


def embed_sharded(graph: SparseMatrix, feature_dim: int = 256, num_iterations: int = 40, propagation: str = "left",
                  normalization: str = "l2", seed: int = 0, initial_embeddings: Optional[np.ndarray] = None,
                  residual_weight: float = 0.0, convergence_threshold: float = 0.0, whiten: bool = True,
                  group=None, backend=None) -> np.ndarray:
    """embed() over all ranks of the default process group; every rank passes the same host graph and gets the
    full result.  Same dispatch as embed(): whiten=False + l2 -> Rust fast-path semantics."""
    import torch.distributed as dist
    if propagation not in _lib.MARKOV:
        raise ValueError(f"Unknown propagation type: '{propagation}'. Use 'left' or 'symmetric'.")
    norms = {"l2": _lib.NORM_L2_NUMPY, "l1": _lib.NORM_L1_NUMPY, "none": _lib.NORM_NONE}
    if normalization not in norms:
        raise ValueError(f"Unknown normalization method: {normalization}. Use 'l2', 'l1', 'spectral', or 'none'.")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    d = feature_dim if initial_embeddings is None else initial_embeddings.shape[1]
    rust = initial_embeddings is None and normalization == "l2" and not whiten
    norm = _lib.NORM_L2_RUST if rust else norms[normalization]
    from . import colsharded
    if backend is None and colsharded.eligible(d, world):
        # column-sharded SpMM with the transposes fused into the kernels' epilogues (colsharded.py): no all-gather
        cem = colsharded.ColumnShardedEmbedder(graph, d, rank, world, group=group)
        try:
            if cem.pipeline_eligible(num_iterations, norm, residual_weight, convergence_threshold, whiten):
                cem.run_pipelined(num_iterations, _lib.MARKOV[propagation], seed, initial_embeddings)
            else:
                cem.run(num_iterations, _lib.MARKOV[propagation], norm, seed, initial_embeddings, residual_weight,
                        convergence_threshold, whiten, rust_semantics=rust)
            return cem.result()
        finally:
            cem.close()
    shard = Shard.from_matrix(graph, rank, world)
    em = ShardedEmbedder(shard, d, backend=backend, group=group)
    if em.pipeline_eligible(num_iterations, norm, residual_weight, convergence_threshold, whiten):
        em.run_pipelined(num_iterations, _lib.MARKOV[propagation], seed, initial_embeddings)
    else:
        em.run(num_iterations, _lib.MARKOV[propagation], norm, seed, initial_embeddings, residual_weight,
               convergence_threshold, whiten, rust_semantics=rust)
    return em.result()


# ------------------------------------------------------------------------------------------------ bench (N > 1)
class _Timers:
    def __init__(self, torch):
        self.torch, self.ev = torch, {}

    def start(self, name):
        a = self.torch.cuda.Event(enable_timing=True)
        a.record()
        return (name, a)

    def stop(self, tok):
        b = self.torch.cuda.Event(enable_timing=True)
        b.record()
        self.ev.setdefault(tok[0], []).append((tok[1], b))

    def totals(self):
        self.torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.ev.items()}


def bench(args, w, name, build_host_graph, spmm_bytes, measured_peaks, ClockSampler):
    """bench.py's N>1 leg: strong scaling of the same workload over WORLD_SIZE GPUs (one rank per GPU, NCCL).
    Column-sharded loop (colsharded.py) where the shape allows it -- every rank then holds the whole CSR -- else the
    row-sharded loop with its all-gather (CLEORA_B200_COLSHARD=0 forces the latter)."""
    import json
    import torch
    import torch.distributed as dist
    from . import colsharded, pinned_empty, synth_pairs
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d, iters = w["d"], args.iters
    device_gen = bool(w.get("device_gen"))
    cols = colsharded.eligible(d, world)
    L = _lib.lib()

    def build_device(u_d, v_d):
        """(whole graph | this rank's row shard) on the device from the device-resident pair arrays."""
        if cols:
            return SparseMatrix.from_edge_arrays_device(u_d, v_d, want_sym=False)
        return Shard.from_device_pairs(u_d, v_d, rank, world, want_sym=False)

    if device_gen:
        # every rank regenerates the same pair stream on its own GPU (counter-based generator) and builds what it needs
        # there; nothing of the graph touches the host (cleora_dev_synth_pairs / cleora_dev_graph_from_pairs)
        u_d, v_d = synth_pairs(w["kind"], w["n"], w["e"], w["seed"], w.get("alpha", 0.5))
        obj = build_device(u_d, v_d)
        E = int(w["e"])
    else:
        # rank 0 builds the CSR once; the others map it from /dev/shm
        tag = f"/dev/shm/cleora_b200_{os.environ.get('MASTER_PORT', '0')}_{name}"
        if rank == 0:
            g0, n_edges = build_host_graph(w)
            rowptr, col, left, sym = g0._csr()
            os.makedirs(tag, exist_ok=True)
            for nm, a in (("rowptr", rowptr), ("col", col), ("left", left), ("sym", sym), ("hash", g0.entity_hashes())):
                np.save(os.path.join(tag, nm + ".npy"), a)
            meta = [int(n_edges)]
            del g0
        else:
            meta = [0]
        dist.barrier()
        dist.broadcast_object_list(meta, src=0)
        E = meta[0]
        arrs = {nm: np.load(os.path.join(tag, nm + ".npy"), mmap_mode="r") for nm in ("rowptr", "col", "left", "sym", "hash")}
        if cols:
            obj = SparseMatrix.from_csr(arrs["rowptr"], arrs["col"], arrs["left"], None, None, arrs["hash"])
        else:
            obj = Shard(arrs["rowptr"], arrs["col"], arrs["left"], arrs["sym"], arrs["hash"], rank, world)
        dist.barrier()
        if rank == 0:
            import shutil
            shutil.rmtree(tag, ignore_errors=True)

    if cols:
        graph = obj
        n, nnz = graph.num_entities, graph.num_edges
        em = colsharded.ColumnShardedEmbedder(graph, d, rank, world)
        n_local, block = em.n_local, em.block
    else:
        shard = obj
        if shard.nnz is None:
            nnz_t = torch.tensor([shard.nnz_local], device="cuda", dtype=torch.int64)
            dist.all_reduce(nnz_t)
            shard.nnz = int(nnz_t.item())
        n, nnz = shard.n, shard.nnz
        em = ShardedEmbedder(shard, d)
        graph = shard.graph
        n_local, block = shard.n_local, shard.block
    check(L.cleora_dev_graph_prepare(graph._handle()))
    dist.barrier()
    norm = _lib.NORM_L2_NUMPY if args.whiten else _lib.NORM_L2_RUST
    piped = em.pipeline_eligible(iters, norm, 0.0, 0.0, bool(args.whiten))

    def step(timers=None):
        if piped:
            em.run_pipelined(iters, 0, 0, None, timers=timers)
        else:
            em.run(iters, 0, norm, 0, None, 0.0, 0.0, bool(args.whiten), rust_semantics=not args.whiten, timers=timers)

    def own_rows():
        return em.result_rows[:n_local] if cols else em._own(em.x_full)[:n_local]

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = L.cleora_kernel_launch_count()
    timers = _Timers(torch)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    ev0.record()
    for _ in range(args.steps):
        step(timers)
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                                     # max over ranks
    ms_step = float(ms.item()) / args.steps
    launches = torch.tensor([L.cleora_kernel_launch_count() - launches0], device="cuda")
    dist.all_reduce(launches)
    tot = timers.totals()
    phases = torch.tensor([tot.get(k, 0.0) for k in ("spmm", "stats", "eigh", "apply", "gather")], device="cuda")
    dist.all_reduce(phases, op=dist.ReduceOp.MAX)

    # e2e: inputs in host memory -> device -> loop -> this rank's rows of the result in pinned host memory.  The full
    # result is the concatenation of the ranks' blocks; every rank writes its own block, as a sharded consumer would.
    # Host-built workloads: the rank's CSR (its shard, or the whole CSR in the column-sharded loop) is re-copied into
    # the same device buffers every step.  Device-built workloads: the edge list (pinned host memory on every rank) is
    # uploaded and the device CSR is rebuilt every step.  Timed per rank from a common barrier; the slowest rank counts.
    own = pinned_empty((max(block, 1), d), np.float32)
    own_t = torch.from_numpy(own)
    if device_gen:
        hu, hv = pinned_empty((E,), np.int32), pinned_empty((E,), np.int32)
        torch.from_numpy(hu).copy_(u_d)
        torch.from_numpy(hv).copy_(v_d)
    e2e_t = []
    for i in range(1 + (max(1, args.e2e_steps - 1) if device_gen else args.e2e_steps)):
        if device_gen:                               # free the previous device image before the rebuild
            if cols:
                em.graph = graph = None
            else:
                em.shard = None
                shard.graph = graph = None
            obj = None
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        if device_gen:
            u_d.copy_(torch.from_numpy(hu), non_blocking=True)
            v_d.copy_(torch.from_numpy(hv), non_blocking=True)
            obj = build_device(u_d, v_d)
            if cols:
                em.graph = graph = obj
                em.hash_ptr, _ = em.be.graph_hashes(graph)
                em._rowscale_markov = None
            else:
                shard = obj
                shard.nnz = nnz
                em.shard, em.hash_pad, graph = shard, _DevPtr(shard.hash_ptr), shard.graph
                if hasattr(em, "_pl"):
                    em._pl["markov"] = None
        else:
            check(L.cleora_graph_refresh_device(graph._handle(), em.be.stream()))
        step()
        own_t[:n_local].copy_(own_rows(), non_blocking=True)
        torch.cuda.synchronize()
        if i > 0:
            e2e_t.append(time.perf_counter() - t0)
    res = own[:n_local]
    e2e = torch.tensor([sum(e2e_t) / len(e2e_t)], device="cuda")
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    clk = clocks.stop() if rank == 0 else None
    if rank == 0:
        peak, peak_src = measured_peaks()
        spmm_ms = float(phases[0].item()) / (iters * args.steps)
        if cols:        # per GPU: the whole CSR once + one column slice of the gathered rows + the slice of the product
            ds = d // world
            b_local = nnz * (8 + 4 * ds) + 8 * (n + 1) + 4 * n * ds
            h2d = 8 * E if device_gen else 8 * (n + 1) + 8 * nnz + 8 * n
        else:
            b_local = spmm_bytes(n, nnz, d) / world
            h2d = 8 * E if device_gen else 8 * (shard.n_local + 1) + 8 * shard.nnz_local + 8 * n
        achieved = b_local / (spmm_ms * 1e-3) / 1e9
        value = E * iters / (ms_step * 1e-3)
        chol = em.be.chol_enabled(d) and bool(args.whiten)
        line = {
            "metric": "edges/sec through the 40-iteration embed() loop", "value": value, "unit": "edges/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic" + (" (generated and ingested on the device, per rank)" if device_gen else ""),
            "config": {"workload": name, "nodes": n, "edges": E, "nnz": nnz, "d": d, "iters": iters,
                       "whiten": bool(args.whiten), "pipeline_whiten": bool(piped),
                       "inner_whitening": ("Cholesky factor, computed redundantly on every rank (no broadcast); PCA eigh "
                                           "on rank 0 + broadcast for the last iterate" if chol
                                           else "PCA eigh on rank 0 + broadcast every iteration"),
                       "parallelism": (f"column-sharded SpMM x{world} (CSR replicated, {d // world} columns per GPU), dense "
                                       f"stages row-sharded" if cols else f"row-shard x{world} (nnz-balanced)"),
                       "collectives": ("two all-to-all transposes per iteration fused into the K1 / K3 epilogues (peer "
                                       "stores over NVLink); all-reduce of d + d*d f64" if cols else
                                       "NCCL all-gather of X blocks per iteration; all-reduce of d+d*d f64"),
                       "l2_flush": "inputs exceed the 126 MB L2"},
            "nnz_per_s": nnz * iters / (ms_step * 1e-3),
            "e2e": {"value": E * iters / float(e2e.item()), "unit": "edges/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(4 * n_local * d), "ms_per_step": 1e3 * float(e2e.item()),
                    "samples": len(e2e_t),
                    "note": "per-rank bytes; each rank re-copies its CSR host->device (device-generated workloads: uploads "
                            "the edge list and rebuilds its CSR on the device) and downloads its own rows of the result "
                            "every step; X0 comes from the init kernel; max over ranks"},
            "gpu_launches": int(launches.item()),
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "spmm_rows_kernel (K1), per GPU", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": b_local, "ms_per_launch": spmm_ms},
            "phase_ms_per_iter": {k: float(phases[i].item()) / (iters * args.steps)
                                  for i, k in enumerate(("spmm", "stats", "eigh", "apply", "gather"))},
            "phase_note": "max over ranks; eigh (inner: Cholesky kernel) runs on a side stream beside spmm; gather = "
                          + ("the barrier after K1's scattered stores" if cols else "NCCL all-gather beside stats"),
            "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
        assert res.shape == (n_local, d)
    dist.barrier()
    dist.destroy_process_group()
