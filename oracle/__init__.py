"""CPU oracle for the Cleora hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  ``cleora_b200`` never does.

Two halves:

* ``liboracle.so`` (``cleora_oracle.c``) -- C restatement of the reference's Rust path (graph build, XXH64,
  FxHash init, SpMM, L2, ``embed_full``), pinned bit-for-bit by the reference's four insta snapshots
  (``tests/golden/snapshots.npz``; see ``tests/golden/make_golden.py``).
* numpy restatement (below) of the reference's *Python* half of the loop -- ``_normalize``,
  ``whiten_embeddings``, ``_compute_rmse`` and the ``embed()`` driver (``pycleora/__init__.py:51-164,942-976``).
  It is validated against the unmodified reference module imported from ``/root/reference`` in this
  container (``tests/golden/make_golden.py`` writes ``tests/golden/embed_*.npz``; the reference tree does not
  exist on the GPU box).  The reference has NO Python tests, so numpy/LAPACK stages are pinned only by
  outputs of the reference code itself run here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Callable, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile liboracle.so (gcc, seconds)."""
    src = os.path.join(_HERE, "cleora_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64]
        L.orc_init_value.restype = C.c_float
        L.orc_init_value.argtypes = [C.c_uint64, C.c_uint64, C.c_int64]
        L.orc_init_matrix.argtypes = [_u64p, C.c_int64, C.c_int64, C.c_int64, _f32p]
        L.orc_spmm.argtypes = [C.c_int64, _i64p, _u32p, _f32p, _f32p, C.c_int64, _f32p]
        L.orc_l2_normalize_inplace.argtypes = [_f32p, C.c_int64, C.c_int64]
        L.orc_embed_full.argtypes = [C.c_int64, _i64p, _u32p, _f32p, _f32p, _f32p, C.c_int64, C.c_int64, C.c_float]
        L.orc_embed_full_convergence.restype = C.c_int64
        L.orc_embed_full_convergence.argtypes = [C.c_int64, _i64p, _u32p, _f32p, _f32p, _f32p, C.c_int64,
                                                 C.c_int64, C.c_float, C.c_float]
        L.orc_graph_build.restype = C.c_void_p
        L.orc_graph_build.argtypes = [C.c_char_p, _i64p, C.c_int64, C.c_char_p, C.c_int64, C.c_char_p, C.c_size_t]
        L.orc_graph_free.argtypes = [C.c_void_p]
        L.orc_graph_from_pairs.restype = C.c_void_p
        L.orc_graph_from_pairs.argtypes = [_u32p, _u32p, C.c_int64, C.c_char_p]
        for name, rt in [("n", C.c_int64), ("nnz", C.c_int64), ("rowptr", C.c_void_p), ("col", C.c_void_p),
                         ("left", C.c_void_p), ("sym", C.c_void_p), ("row_sum", C.c_void_p),
                         ("hash", C.c_void_p), ("column_id", C.c_void_p)]:
            f = getattr(L, "orc_graph_" + name)
            f.restype = rt
            f.argtypes = [C.c_void_p]
        L.orc_graph_id.restype = C.c_char_p
        L.orc_graph_id.argtypes = [C.c_void_p, C.c_int64]
        L.orc_graph_col_name.restype = C.c_char_p
        L.orc_graph_col_name.argtypes = [C.c_void_p, C.c_int]
        L.orc_graph_col_id.restype = C.c_int
        L.orc_graph_col_id.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


# ------------------------------------------------------------------------------------------------ hashing / init
def xxh64(data: bytes, seed: int = 0) -> int:
    return int(lib().orc_xxh64(data, len(data), seed))


def init_value(col: int, hsh: int, seed: int = 0) -> float:
    return float(lib().orc_init_value(col, hsh, seed))


def init_matrix(hashes: np.ndarray, d: int, seed: int = 0) -> np.ndarray:
    hashes = np.ascontiguousarray(hashes, dtype=np.uint64)
    out = np.empty((hashes.shape[0], d), dtype=np.float32)
    lib().orc_init_matrix(hashes, hashes.shape[0], d, seed, out)
    return out


# ------------------------------------------------------------------------------------------------ graph
class OracleGraph:
    """CSR + entity tables as the reference's SparseMatrix holds them (src/sparse_matrix.rs:56-78)."""

    def __init__(self, rowptr, col, left, sym, row_sum, hashes, column_ids, entity_ids,
                 col_a_name="", col_b_name="", col_a_id=0, col_b_id=0):
        self.rowptr = np.ascontiguousarray(rowptr, np.int64)
        self.col = np.ascontiguousarray(col, np.uint32)
        self.left = np.ascontiguousarray(left, np.float32)
        self.sym = np.ascontiguousarray(sym, np.float32)
        self.row_sum = np.ascontiguousarray(row_sum, np.float32)
        self.hashes = np.ascontiguousarray(hashes, np.uint64)
        self.column_ids = np.ascontiguousarray(column_ids, np.uint8)
        self.entity_ids = entity_ids
        self.col_a_name, self.col_b_name = col_a_name, col_b_name
        self.col_a_id, self.col_b_id = col_a_id, col_b_id

    @property
    def n(self) -> int:
        return self.rowptr.shape[0] - 1

    @property
    def nnz(self) -> int:
        return self.col.shape[0]

    def values(self, propagation: str = "left") -> np.ndarray:
        if propagation not in ("left", "symmetric"):
            raise ValueError(f"Unknown propagation '{propagation}'. Use 'left' or 'symmetric'.")
        return self.left if propagation == "left" else self.sym


def build_graph(lines, columns: str, hyperedge_trim_n: int = 16) -> OracleGraph:
    """SparseMatrix.from_iterator semantics (src/lib.rs:104-135 -> pipeline.rs:24-42)."""
    enc = [s.encode("utf-8") for s in lines]
    offsets = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        offsets[1:] = np.cumsum([len(b) for b in enc])
    buf = b"".join(enc)
    err = C.create_string_buffer(512)
    L = lib()
    h = L.orc_graph_build(buf, offsets, len(enc), columns.encode(), hyperedge_trim_n, err, 512)
    if not h:
        raise ValueError(err.value.decode())
    try:
        n, nnz = L.orc_graph_n(h), L.orc_graph_nnz(h)

        def arr(ptr, count, dt):
            if count == 0:
                return np.zeros(0, dt)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(count,))
            return a.copy()

        g = OracleGraph(
            arr(L.orc_graph_rowptr(h), n + 1, np.int64), arr(L.orc_graph_col(h), nnz, np.uint32),
            arr(L.orc_graph_left(h), nnz, np.float32), arr(L.orc_graph_sym(h), nnz, np.float32),
            arr(L.orc_graph_row_sum(h), n, np.float32), arr(L.orc_graph_hash(h), n, np.uint64),
            arr(L.orc_graph_column_id(h), n, np.uint8),
            [L.orc_graph_id(h, i).decode("utf-8") for i in range(n)],
            L.orc_graph_col_name(h, 0).decode(), L.orc_graph_col_name(h, 1).decode(),
            L.orc_graph_col_id(h, 0), L.orc_graph_col_id(h, 1))
    finally:
        L.orc_graph_free(h)
    return g


def graph_from_pairs(src, dst, column_name: str = "node") -> OracleGraph:
    """The graph ``build_graph((f"{u} {v}" ...), "complex::reflexive::name")`` builds, from the id arrays
    (``orc_graph_from_pairs``).  ``entity_ids`` is None (the ids are the decimal strings, not materialised)."""
    u = np.ascontiguousarray(src, np.uint32)
    v = np.ascontiguousarray(dst, np.uint32)
    assert u.shape == v.shape and u.ndim == 1
    L = lib()
    h = L.orc_graph_from_pairs(u, v, u.shape[0], column_name.encode())
    try:
        n, nnz = L.orc_graph_n(h), L.orc_graph_nnz(h)

        def arr(ptr, count, dt):
            if count == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(count,)).copy()

        return OracleGraph(
            arr(L.orc_graph_rowptr(h), n + 1, np.int64) if n else np.zeros(1, np.int64),
            arr(L.orc_graph_col(h), nnz, np.uint32), arr(L.orc_graph_left(h), nnz, np.float32),
            arr(L.orc_graph_sym(h), nnz, np.float32), arr(L.orc_graph_row_sum(h), n, np.float32),
            arr(L.orc_graph_hash(h), n, np.uint64), arr(L.orc_graph_column_id(h), n, np.uint8), None,
            column_name, column_name, 0, 1)
    finally:
        L.orc_graph_free(h)


# ------------------------------------------------------------------------------------------------ Rust-side compute
def spmm(g: OracleGraph, x: np.ndarray, propagation: str = "left") -> np.ndarray:
    """NdArrayMatrix::multiply (src/embedding.rs:15-39)."""
    x = np.ascontiguousarray(x, np.float32)
    assert x.shape[0] == g.n
    out = np.zeros_like(x)
    lib().orc_spmm(g.n, g.rowptr, g.col, g.values(propagation), x, x.shape[1], out)
    return out


def l2_normalize(x: np.ndarray) -> np.ndarray:
    """SparseMatrix.l2_normalize (src/lib.rs:414-424 -> embedding.rs:88-104)."""
    out = np.array(x, dtype=np.float32, order="C", copy=True)
    lib().orc_l2_normalize_inplace(out, out.shape[0], out.shape[1])
    return out


def embed_fast(g: OracleGraph, d: int, iters: int, propagation: str = "left", seed: int = 0,
               residual_weight: float = 0.0, x0: Optional[np.ndarray] = None) -> np.ndarray:
    """SparseMatrix.embed_fast (src/lib.rs:320-364 -> embedding.rs:106-136)."""
    x = init_matrix(g.hashes, d, seed) if x0 is None else np.array(x0, np.float32, order="C", copy=True)
    tmp = np.zeros_like(x)
    lib().orc_embed_full(g.n, g.rowptr, g.col, g.values(propagation), x, tmp, d, iters, residual_weight)
    return x


def embed_fast_convergence(g: OracleGraph, d: int, max_iters: int, propagation: str = "left", seed: int = 0,
                           residual_weight: float = 0.0, convergence_threshold: float = 0.0
                           ) -> Tuple[np.ndarray, int]:
    """SparseMatrix.embed_fast_convergence (src/lib.rs:366-412 -> embedding.rs:138-188)."""
    x = init_matrix(g.hashes, d, seed)
    tmp = np.zeros_like(x)
    it = lib().orc_embed_full_convergence(g.n, g.rowptr, g.col, g.values(propagation), x, tmp, d, max_iters,
                                          residual_weight, convergence_threshold)
    return x, int(it)


# ------------------------------------------------------------------------------------------------ Python-side compute
def normalize(x: np.ndarray, method: str = "l2") -> np.ndarray:
    """_normalize (pycleora/__init__.py:942-960), methods on the path: l2 / l1 / none."""
    if method == "l2":
        norms = np.maximum(np.linalg.norm(x, ord=2, axis=-1, keepdims=True), 1e-10)
        return x / norms
    if method == "l1":
        norms = np.maximum(np.linalg.norm(x, ord=1, axis=-1, keepdims=True), 1e-10)
        return x / norms
    if method == "none":
        return x
    raise ValueError(f"Unknown normalization method: {method}. Use 'l2', 'l1', 'spectral', or 'none'.")


def whiten_stats(x: np.ndarray, chunk: int = 50000) -> Tuple[np.ndarray, np.ndarray]:
    """Mean (f64) and centred covariance (f64) exactly as pycleora/__init__.py:136-143 computes them."""
    n, d = x.shape
    mean = x.mean(axis=0, dtype=np.float64)
    cov = np.zeros((d, d), dtype=np.float64)
    for i in range(0, n, chunk):
        block = x[i:min(i + chunk, n)].astype(np.float64) - mean
        cov += block.T @ block
    cov *= 1.0 / (n - 1)
    return mean, cov


def whiten_transform(cov: np.ndarray, n_components: Optional[int] = None) -> np.ndarray:
    """eigh -> sort descending -> scale -> f32 transform (pycleora/__init__.py:145-156)."""
    eigenvalues, eigenvectors = np.linalg.eigh(cov)
    idx = np.argsort(eigenvalues)[::-1]
    eigenvalues, eigenvectors = eigenvalues[idx], eigenvectors[:, idx]
    if n_components is not None:
        eigenvalues, eigenvectors = eigenvalues[:n_components], eigenvectors[:, :n_components]
    scale = 1.0 / np.sqrt(np.maximum(eigenvalues, 1e-10))
    return (eigenvectors * scale).astype(np.float32)


def whiten_apply(x: np.ndarray, mean: np.ndarray, transform: np.ndarray, chunk: int = 50000) -> np.ndarray:
    """(X - mean_f32) @ T in f32 chunks (pycleora/__init__.py:157-163)."""
    n = x.shape[0]
    mean_f32 = mean.astype(np.float32)
    out = np.empty((n, transform.shape[1]), dtype=np.float32)
    for i in range(0, n, chunk):
        end = min(i + chunk, n)
        np.dot(x[i:end] - mean_f32, transform, out=out[i:end])
    return out


def whiten_embeddings(x: np.ndarray, n_components: Optional[int] = None) -> np.ndarray:
    """whiten_embeddings (pycleora/__init__.py:130-164) -- PCA whitening."""
    if x.shape[0] <= 1:
        return x.copy()
    mean, cov = whiten_stats(x)
    return whiten_apply(x, mean, whiten_transform(cov, n_components))


def compute_rmse(cur: np.ndarray, prev: np.ndarray) -> float:
    """_compute_rmse (pycleora/__init__.py:974-976)."""
    diff = cur.astype(np.float64, copy=False) - prev.astype(np.float64, copy=False)
    return float(np.sqrt(np.mean(diff * diff)))


def embed(g: OracleGraph, feature_dim: int = 256, num_iterations: int = 40, propagation: str = "left",
          normalization: str = "l2", seed: int = 0, initial_embeddings: Optional[np.ndarray] = None,
          callback: Optional[Callable[[int, np.ndarray], None]] = None, residual_weight: float = 0.0,
          convergence_threshold: float = 0.0, whiten: bool = True) -> np.ndarray:
    """embed() (pycleora/__init__.py:51-127): fast path when whiten=False (Rust loop), else the Python loop
    propagate -> residual -> _normalize -> whiten_embeddings -> callback -> rmse."""
    if propagation not in ("left", "symmetric"):
        raise ValueError(f"Unknown propagation type: '{propagation}'. Use 'left' or 'symmetric'.")
    fast = initial_embeddings is None and callback is None and normalization == "l2" and not whiten
    if fast:
        if convergence_threshold > 0:
            return embed_fast_convergence(g, feature_dim, num_iterations, propagation, seed, residual_weight,
                                          convergence_threshold)[0]
        return embed_fast(g, feature_dim, num_iterations, propagation, seed, residual_weight)
    if initial_embeddings is not None:
        emb = initial_embeddings.astype(np.float32)
        if emb.shape[0] != g.n:
            raise ValueError(f"initial_embeddings has {emb.shape[0]} rows but graph has {g.n} entities")
    else:
        emb = init_matrix(g.hashes, feature_dim, seed)
    for i in range(num_iterations):
        prev = emb
        emb = spmm(g, emb, propagation)
        if residual_weight > 0:
            emb = (1 - residual_weight) * emb + residual_weight * prev
        emb = normalize(emb, normalization)
        if whiten:
            emb = whiten_embeddings(emb)
        if callback is not None:
            callback(i, emb)
        if convergence_threshold > 0 and i > 0 and compute_rmse(emb, prev) < convergence_threshold:
            break
    return emb
