"""``SparseMatrix`` -- host-side mirror of the reference's pyo3 class (``src/lib.rs:84-476``,
stub ``pycleora/pycleora.pyi:7-36``) bound to libcleora_b200's C ABI.

Same names, defaults, argument meaning and exception classes as the reference, so code written against
``pycleora.pycleora.SparseMatrix`` runs unchanged.  All compute goes through CUDA kernels; ``num_workers`` is
accepted and ignored.  The reference-side (Rust) binding of the same ABI is shown in INTEGRATION.md.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import check, ptr


_DEVICES: List[int] = []


def set_devices(devices: Optional[Iterable[int]]) -> None:
    """GPUs that ``embed()`` / ``embed_fast*()`` / ``embed_device()`` spread one call over, from this one process
    (include/cleora_b200.h: cleora_embed_multi -- column-sharded SpMM, peer stores over NVLink, no torchrun).  ``None``
    or a single device restores the one-GPU path; the environment variable CLEORA_B200_DEVICES="0,1,2,3" sets the
    initial list.  Shapes the multi-GPU loop has no slice kernels for (``d % len(devices) != 0`` ...) run on one GPU."""
    global _DEVICES
    _DEVICES = [int(v) for v in devices] if devices is not None else []


def get_devices() -> List[int]:
    return list(_DEVICES)


def _multi_devices(d: int):
    """ctypes int array of the configured devices if this feature dimension can be column-sharded over them."""
    if len(_DEVICES) < 2 or not _lib.lib().cleora_embed_multi_supported(int(d), len(_DEVICES)):
        return None
    return (C.c_int * len(_DEVICES))(*_DEVICES)


import os as _os  # noqa: E402
if _os.environ.get("CLEORA_B200_DEVICES"):
    set_devices(int(v) for v in _os.environ["CLEORA_B200_DEVICES"].split(",") if v.strip())


def _as_matrix(x, name="x") -> np.ndarray:
    """pyo3 accepts only a float32 ndarray with ndim == 2 (`&PyArray2<f32>`); anything else is a TypeError."""
    if not isinstance(x, np.ndarray) or x.dtype != np.float32 or x.ndim != 2:
        raise TypeError(f"argument '{name}': expected a 2-D numpy array of float32")
    return np.ascontiguousarray(x)      # the reference panics on non-contiguous rows; we copy instead


def _pack_strings(items: List[bytes]) -> Tuple[bytes, np.ndarray]:
    offsets = np.zeros(len(items) + 1, dtype=np.int64)
    if items:
        np.cumsum([len(b) for b in items], out=offsets[1:])
    return b"".join(items), offsets


class SparseMatrix:
    """CSR Markov operator of a (hyper)graph, resident in HBM once first used."""

    def __init__(self, *args):
        # src/lib.rs:440-461: zero-arg construction only (used by pickle)
        if args:
            raise ValueError("SparseMatrix cannot be constructed directly. Use SparseMatrix.from_files() or "
                             "SparseMatrix.from_iterator().")
        self._h = None
        self._ids_cache = None
        self._csr_cache = None

    # ------------------------------------------------------------------------------------------ lifetime
    @classmethod
    def _adopt(cls, handle) -> "SparseMatrix":
        sm = cls()
        sm._h = handle
        return sm

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().cleora_graph_destroy(h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass

    def _handle(self):
        if not self._h:
            # an empty SparseMatrix() behaves like the reference's empty struct
            h = C.c_void_p()
            z64 = np.zeros(1, np.int64)
            check(_lib.lib().cleora_graph_from_csr(ptr(z64, _lib.c_i64p), None, None, None, None, None, 0, 0, 0,
                                                   C.byref(h)))
            self._h = h
        return self._h

    # ------------------------------------------------------------------------------------------ constructors
    @staticmethod
    def from_iterator(hyperedges: Iterable[str], columns: str, hyperedge_trim_n: int = 16,
                      num_workers: Optional[int] = None) -> "SparseMatrix":
        """src/lib.rs:104-135.  Every element must be a ``str``."""
        enc = []
        for line in hyperedges:
            if not isinstance(line, str):
                raise ValueError("Iterator elements must be strings")
            enc.append(line.encode("utf-8"))
        buf, offsets = _pack_strings(enc)
        h = C.c_void_p()
        check(_lib.lib().cleora_graph_from_lines(buf, ptr(offsets, _lib.c_i64p), len(enc), columns.encode("utf-8"),
                                                 int(hyperedge_trim_n), C.byref(h)))
        return SparseMatrix._adopt(h)

    @staticmethod
    def from_files(filepaths: List[str], columns: str, hyperedge_trim_n: int = 16,
                   num_workers: Optional[int] = None) -> "SparseMatrix":
        """src/lib.rs:137-173 (.tsv/.csv/.txt only; files are read sequentially => deterministic entity order)."""
        paths = [p.encode("utf-8") for p in filepaths]
        arr = (C.c_char_p * len(paths))(*paths)
        h = C.c_void_p()
        check(_lib.lib().cleora_graph_from_files(arr, len(paths), columns.encode("utf-8"), int(hyperedge_trim_n),
                                                 C.byref(h)))
        return SparseMatrix._adopt(h)

    @staticmethod
    def from_edge_arrays(src: np.ndarray, dst: np.ndarray, column_name: str = "node") -> "SparseMatrix":
        """Integer ingest: the graph ``from_iterator((f"{u} {v}" ...), "complex::reflexive::name")``
        builds, without going through strings."""
        u = np.ascontiguousarray(src, dtype=np.uint32)
        v = np.ascontiguousarray(dst, dtype=np.uint32)
        if u.shape != v.shape or u.ndim != 1:
            raise ValueError("src and dst must be 1-D arrays of the same length")
        h = C.c_void_p()
        check(_lib.lib().cleora_graph_from_pairs(ptr(u, _lib.c_u32p), ptr(v, _lib.c_u32p), u.shape[0],
                                                 column_name.encode("utf-8"), C.byref(h)))
        return SparseMatrix._adopt(h)

    @staticmethod
    def from_hyperedge_arrays(members: np.ndarray, offsets: np.ndarray, columns: str = "complex::reflexive::node",
                              hyperedge_trim_n: int = 16) -> "SparseMatrix":
        """The graph ``from_iterator((" ".join(map(str, line)) for line in hyperedges), columns)`` builds, from a flat
        member array and line offsets (hyperedge i = ``members[offsets[i]:offsets[i+1]]``), without Python strings."""
        m = np.ascontiguousarray(members, dtype=np.uint32)
        o = np.ascontiguousarray(offsets, dtype=np.int64)
        if o.ndim != 1 or o.shape[0] < 1 or m.ndim != 1 or int(o[-1]) != m.shape[0]:
            raise ValueError("offsets must have one entry per hyperedge plus one and end at len(members)")
        h = C.c_void_p()
        check(_lib.lib().cleora_graph_from_hyperedges(ptr(m, _lib.c_u32p), ptr(o, _lib.c_i64p), o.shape[0] - 1,
                                                      columns.encode("utf-8"), int(hyperedge_trim_n), C.byref(h)))
        return SparseMatrix._adopt(h)

    @staticmethod
    def from_edge_arrays_device(src, dst, column_name: str = "node", shard_rank: int = 0, shard_world: int = 1,
                                want_sym: bool = True, stream: int = 0) -> "SparseMatrix":
        """``from_edge_arrays`` on the GPU: ``src`` / ``dst`` are DEVICE arrays of 32-bit ids on the current device
        (anything with ``data_ptr()`` and ``numel()``, e.g. torch int32 / uint32 tensors).  The CSR is built and kept in
        HBM (``include/cleora_b200.h``: cleora_dev_graph_from_pairs); with ``shard_world > 1`` only the row block
        ``shard_rank`` is built, in the padded gathered layout of ``cleora_b200.sharded``.  ``.shard_bounds`` holds the
        row partition."""
        n_pairs = int(src.numel())
        if int(dst.numel()) != n_pairs:
            raise ValueError("src and dst must have the same length")
        h = C.c_void_p()
        bounds = np.zeros(int(shard_world) + 1, np.int64)
        check(_lib.lib().cleora_dev_graph_from_pairs(src.data_ptr(), dst.data_ptr(), n_pairs, column_name.encode("utf-8"),
                                                     int(shard_rank), int(shard_world), 1 if want_sym else 0, stream,
                                                     C.byref(h), ptr(bounds, _lib.c_i64p)))
        sm = SparseMatrix._adopt(h)
        sm.shard_bounds = bounds
        return sm

    @staticmethod
    def from_csr(rowptr, col, val_left, val_sym=None, row_sum=None, entity_hash=None, n_cols=None,
                 row_offset: int = 0) -> "SparseMatrix":
        """Adopt a prebuilt CSR (a full graph or a row shard of one)."""
        rowptr = np.ascontiguousarray(rowptr, np.int64)
        col = np.ascontiguousarray(col, np.uint32)
        val_left = np.ascontiguousarray(val_left, np.float32)
        n_rows = rowptr.shape[0] - 1
        keep = [np.ascontiguousarray(a, dt) if a is not None else None
                for a, dt in ((val_sym, np.float32), (row_sum, np.float32), (entity_hash, np.uint64))]
        h = C.c_void_p()
        check(_lib.lib().cleora_graph_from_csr(
            ptr(rowptr, _lib.c_i64p), ptr(col, _lib.c_u32p), ptr(val_left, _lib.c_f32p),
            None if keep[0] is None else ptr(keep[0], _lib.c_f32p),
            None if keep[1] is None else ptr(keep[1], _lib.c_f32p),
            None if keep[2] is None else ptr(keep[2], _lib.c_u64p),
            n_rows, n_rows if n_cols is None else int(n_cols), int(row_offset), C.byref(h)))
        return SparseMatrix._adopt(h)

    # ------------------------------------------------------------------------------------------ introspection
    @property
    def num_entities(self) -> int:
        return int(_lib.lib().cleora_graph_num_entities(self._handle()))

    @property
    def num_edges(self) -> int:
        return int(_lib.lib().cleora_graph_num_edges(self._handle()))

    def __len__(self) -> int:
        return self.num_entities

    def __repr__(self) -> str:
        L, h = _lib.lib(), self._handle()
        return "SparseMatrix(entities={}, edges={}, columns=('{}', '{}'))".format(
            self.num_entities, self.num_edges, L.cleora_graph_col_name(h, 0).decode(),
            L.cleora_graph_col_name(h, 1).decode())

    @property
    def entity_ids(self) -> List[str]:
        """#[pyo3(get, set)] (src/sparse_matrix.rs:60-61).  The list is cached (the reference clones it on every
        access); a fresh shallow copy is returned so callers cannot mutate the cache."""
        if self._ids_cache is None:
            L, h = _lib.lib(), self._handle()
            n = self.num_entities
            nbytes = int(L.cleora_graph_entity_ids_nbytes(h))
            buf = C.create_string_buffer(max(nbytes, 1))
            offsets = np.zeros(n + 1, np.int64)
            check(L.cleora_graph_copy_entity_ids(h, buf, ptr(offsets, _lib.c_i64p)))
            raw = buf.raw
            self._ids_cache = [raw[offsets[i]:offsets[i + 1]].decode("utf-8") for i in range(n)]
        return list(self._ids_cache)

    @entity_ids.setter
    def entity_ids(self, ids: List[str]) -> None:
        buf, offsets = _pack_strings([s.encode("utf-8") for s in ids])
        check(_lib.lib().cleora_graph_set_entity_ids(self._handle(), buf, ptr(offsets, _lib.c_i64p), len(ids)))
        self._ids_cache = list(ids)

    @property
    def entity_degrees(self) -> np.ndarray:
        out = np.empty(self.num_entities, np.float32)
        check(_lib.lib().cleora_graph_copy_row_sums(self._handle(), ptr(out, _lib.c_f32p)))
        return out

    def _csr(self):
        if self._csr_cache is None:
            n, nnz = self.num_entities, self.num_edges
            rowptr, col = np.empty(n + 1, np.int64), np.empty(nnz, np.uint32)
            left, sym = np.empty(nnz, np.float32), np.empty(nnz, np.float32)
            L, h = _lib.lib(), self._handle()
            rc = L.cleora_graph_copy_csr(h, ptr(rowptr, _lib.c_i64p), ptr(col, _lib.c_u32p), ptr(left, _lib.c_f32p),
                                         ptr(sym, _lib.c_f32p))
            if rc == _lib.ERR_VALUE:             # built without symmetric values (adopted CSR / want_sym=False)
                sym = None
                rc = L.cleora_graph_copy_csr(h, ptr(rowptr, _lib.c_i64p), ptr(col, _lib.c_u32p), ptr(left, _lib.c_f32p), None)
            check(rc)
            self._csr_cache = (rowptr, col, left, sym)
        return self._csr_cache

    def entity_hashes(self) -> np.ndarray:
        out = np.empty(self.num_entities, np.uint64)
        check(_lib.lib().cleora_graph_copy_entity_hashes(self._handle(), ptr(out, _lib.c_u64p)))
        return out

    def get_entity_column_mask(self, column_name: str) -> np.ndarray:
        """src/lib.rs:175-198, including its quirk: the name -> id map is built from (col_a, col_b) in that order,
        so for a reflexive column (both names equal) the LAST id wins."""
        L, h = _lib.lib(), self._handle()
        names = {}
        for which in (0, 1):
            names[L.cleora_graph_col_name(h, which).decode()] = L.cleora_graph_col_id(h, which)
        if column_name not in names:
            a, b = L.cleora_graph_col_name(h, 0).decode(), L.cleora_graph_col_name(h, 1).decode()
            raise ValueError(f"Column name '{column_name}' not found. Available: '{a}', '{b}'")
        ids = np.empty(self.num_entities, np.uint8)
        check(L.cleora_graph_copy_column_ids(h, ptr(ids, _lib.c_u8p)))
        return ids == names[column_name]

    def get_entity_index(self, entity_id: str) -> int:
        b = entity_id.encode("utf-8")
        ix = int(_lib.lib().cleora_graph_find_entity(self._handle(), b, len(b)))
        if ix < 0:
            raise ValueError(f"Entity '{entity_id}' not found")
        return ix

    def get_entity_indices(self, entity_ids: List[str]) -> List[int]:
        return [self.get_entity_index(e) for e in entity_ids]

    def get_neighbors(self, entity_id: str) -> List[Tuple[str, float]]:
        """src/lib.rs:302-318: (neighbour id, left Markov value) in column order."""
        ix = self.get_entity_index(entity_id)
        rowptr, col, left, _ = self._csr()
        ids = self._ids_cache if self._ids_cache is not None else self.entity_ids
        return [(ids[int(c)], float(v)) for c, v in zip(col[rowptr[ix]:rowptr[ix + 1]], left[rowptr[ix]:rowptr[ix + 1]])]

    def to_sparse_csr(self, markov_type: Optional[str] = None):
        """src/lib.rs:254-300: COO triplets (rows u32, cols u32, vals f32, n, n) despite the name."""
        mt = "left" if markov_type is None else markov_type
        if mt not in ("left", "symmetric"):
            raise ValueError(f"Unknown markov_type '{mt}'. Use 'left' or 'symmetric'.")
        rowptr, col, left, sym = self._csr()
        n = self.num_entities
        rows = np.repeat(np.arange(n, dtype=np.uint32), np.diff(rowptr))
        return rows, col.copy(), (sym if mt == "symmetric" else left).copy(), n, n

    # ------------------------------------------------------------------------------------------ hot path
    def _propagate(self, x, markov: int) -> np.ndarray:
        x = _as_matrix(x)
        out = np.empty((self.num_entities, x.shape[1]), np.float32)
        check(_lib.lib().cleora_markov_propagate(self._handle(), ptr(x, _lib.c_f32p), x.shape[0], x.shape[1], markov,
                                                 ptr(out, _lib.c_f32p)))
        return out

    def left_markov_propagate(self, x: np.ndarray, num_workers: Optional[int] = None) -> np.ndarray:
        """src/lib.rs:86-93."""
        return self._propagate(x, 0)

    def symmetric_markov_propagate(self, x: np.ndarray, num_workers: Optional[int] = None) -> np.ndarray:
        """src/lib.rs:95-102."""
        return self._propagate(x, 1)

    def initialize_deterministically(self, feature_dim: int, seed: int = 0) -> np.ndarray:
        """src/lib.rs:242-252."""
        out = np.empty((self.num_entities, int(feature_dim)), np.float32)
        check(_lib.lib().cleora_initialize_deterministically(self._handle(), int(feature_dim), int(seed),
                                                             ptr(out, _lib.c_f32p)))
        return out

    @staticmethod
    def _markov_code(propagation: str) -> int:
        if propagation not in _lib.MARKOV:
            raise ValueError(f"Unknown propagation '{propagation}'. Use 'left' or 'symmetric'.")   # src/lib.rs:335-338
        return _lib.MARKOV[propagation]

    def embed_fast(self, feature_dim: int, num_iterations: int, propagation: str = "left", seed: int = 0,
                   residual_weight: float = 0.0, num_workers: Optional[int] = None) -> np.ndarray:
        """src/lib.rs:320-364: init + T x (SpMM -> residual -> L2) with X resident in HBM."""
        m = self._markov_code(propagation)
        out = np.empty((self.num_entities, int(feature_dim)), np.float32)
        devs = _multi_devices(int(feature_dim))
        if devs is not None:
            check(_lib.lib().cleora_embed_multi(self._handle(), devs, len(devs), None, int(feature_dim), int(num_iterations), m,
                                                int(seed), float(np.float32(residual_weight)), 0.0, _lib.NORM_L2_RUST, 0,
                                                ptr(out, _lib.c_f32p), None))
            return out
        check(_lib.lib().cleora_embed_fast(self._handle(), int(feature_dim), int(num_iterations), m, int(seed),
                                           float(residual_weight), ptr(out, _lib.c_f32p)))
        return out

    def embed_fast_convergence(self, feature_dim: int, max_iterations: int, propagation: str = "left", seed: int = 0,
                               residual_weight: float = 0.0, convergence_threshold: float = 0.0,
                               num_workers: Optional[int] = None) -> Tuple[np.ndarray, int]:
        """src/lib.rs:366-412."""
        m = self._markov_code(propagation)
        out = np.empty((self.num_entities, int(feature_dim)), np.float32)
        done = C.c_int64(0)
        devs = _multi_devices(int(feature_dim))
        if devs is not None:
            check(_lib.lib().cleora_embed_multi(self._handle(), devs, len(devs), None, int(feature_dim), int(max_iterations), m,
                                                int(seed), float(np.float32(residual_weight)),
                                                float(np.float32(convergence_threshold)), _lib.NORM_L2_RUST, 0,
                                                ptr(out, _lib.c_f32p), C.byref(done)))
            return out, int(done.value)
        check(_lib.lib().cleora_embed_fast_convergence(self._handle(), int(feature_dim), int(max_iterations), m,
                                                       int(seed), float(residual_weight),
                                                       float(convergence_threshold), ptr(out, _lib.c_f32p),
                                                       C.byref(done)))
        return out, int(done.value)

    def l2_normalize(self, x: np.ndarray, num_workers: Optional[int] = None) -> np.ndarray:
        """src/lib.rs:414-424."""
        x = _as_matrix(x)
        out = np.empty_like(x)
        check(_lib.lib().cleora_l2_normalize(ptr(x, _lib.c_f32p), x.shape[0], x.shape[1], ptr(out, _lib.c_f32p)))
        return out

    def embed_device(self, feature_dim: int, num_iterations: int, propagation: str = "left",
                     normalization: int = _lib.NORM_L2_NUMPY, seed: int = 0,
                     initial_embeddings: Optional[np.ndarray] = None, residual_weight: float = 0.0,
                     convergence_threshold: float = 0.0, whiten: bool = True, out: Optional[np.ndarray] = None,
                     timings: Optional[np.ndarray] = None) -> Tuple[np.ndarray, int]:
        """The whole Python loop of ``embed()`` (pycleora/__init__.py:97-125) in one device-resident call."""
        m = self._markov_code(propagation)
        x0 = None
        d = int(feature_dim)
        if initial_embeddings is not None:
            x0 = np.ascontiguousarray(initial_embeddings, np.float32)
            if x0.ndim != 2 or x0.shape[0] != self.num_entities:
                raise ValueError(f"initial_embeddings has shape {tuple(x0.shape)} but graph has "
                                 f"{self.num_entities} entities")
            d = x0.shape[1]
        if out is None:
            out = np.empty((self.num_entities, d), np.float32)
        elif (not isinstance(out, np.ndarray) or out.dtype != np.float32 or out.shape != (self.num_entities, d)
              or not out.flags["C_CONTIGUOUS"]):
            raise ValueError(f"out must be a C-contiguous float32 array of shape ({self.num_entities}, {d})")
        if timings is not None and (timings.dtype != np.float64 or timings.size < 8):
            raise ValueError("timings must be a float64 array with at least 8 entries")
        done = C.c_int64(0)
        host = _lib.auto_host_eigh(self.num_entities, d, int(num_iterations), int(normalization), bool(whiten),
                                   float(residual_weight), float(convergence_threshold))
        devs = _multi_devices(d) if timings is None else None
        if devs is not None:
            with _lib.host_eigh(host):
                check(_lib.lib().cleora_embed_multi(self._handle(), devs, len(devs), None if x0 is None else ptr(x0, _lib.c_f32p),
                                                    d, int(num_iterations), m, int(seed), float(residual_weight),
                                                    float(convergence_threshold), int(normalization), 1 if whiten else 0,
                                                    ptr(out, _lib.c_f32p), C.byref(done)))
            return out, int(done.value)
        with _lib.host_eigh(host):
            check(_lib.lib().cleora_embed(self._handle(), None if x0 is None else ptr(x0, _lib.c_f32p), d,
                                          int(num_iterations), m, int(seed), float(residual_weight),
                                          float(convergence_threshold), int(normalization), 1 if whiten else 0,
                                          ptr(out, _lib.c_f32p), C.byref(done),
                                          None if timings is None else ptr(timings, _lib.c_f64p)))
        return out, int(done.value)

    # ------------------------------------------------------------------------------------------ pickle (bincode 1.3.3)
    def __getstate__(self) -> bytes:
        """src/lib.rs:463-468: bincode of the struct, field order descriptor, entity_ids, entities, edges, slices,
        column_ids (src/sparse_matrix.rs:56-66); little-endian fixed ints, u64 lengths."""
        L, h = _lib.lib(), self._handle()
        rowptr, col, left, sym = self._csr()
        n, nnz = self.num_entities, self.num_edges
        u64 = lambda v: int(v).to_bytes(8, "little")  # noqa: E731
        string = lambda s: u64(len(s)) + s            # noqa: E731
        out = [bytes([L.cleora_graph_col_id(h, 0)]), string(L.cleora_graph_col_name(h, 0)),
               bytes([L.cleora_graph_col_id(h, 1)]), string(L.cleora_graph_col_name(h, 1)), u64(n)]
        out += [string(s.encode("utf-8")) for s in self.entity_ids]
        out += [u64(n), self.entity_degrees.astype("<f4").tobytes()]
        edges = np.empty(nnz, dtype=[("c", "<u4"), ("l", "<f4"), ("s", "<f4")])
        edges["c"], edges["l"], edges["s"] = col, left, sym
        out += [u64(nnz), edges.tobytes()]
        slices = np.empty((n, 2), "<u8")
        slices[:, 0], slices[:, 1] = rowptr[:-1], rowptr[1:]
        out += [u64(n), slices.tobytes()]
        ids = np.empty(n, np.uint8)
        check(L.cleora_graph_copy_column_ids(h, ptr(ids, _lib.c_u8p)))
        out += [u64(n), ids.tobytes()]
        return b"".join(out)

    def __setstate__(self, state: bytes) -> None:
        """src/lib.rs:470-475."""
        try:
            mv, pos = memoryview(state), 0

            def take(k):
                nonlocal pos
                if pos + k > len(mv):
                    raise ValueError("unexpected end of input")
                b = mv[pos:pos + k]
                pos += k
                return b

            u64 = lambda: int.from_bytes(take(8), "little")  # noqa: E731
            col_a_id = take(1)[0]; name_a = bytes(take(u64()))
            col_b_id = take(1)[0]; name_b = bytes(take(u64()))
            n = u64()
            ids = [bytes(take(u64())).decode("utf-8") for _ in range(n)]
            n_ent = u64(); row_sum = np.frombuffer(take(4 * n_ent), "<f4")
            nnz = u64(); edges = np.frombuffer(take(12 * nnz), dtype=[("c", "<u4"), ("l", "<f4"), ("s", "<f4")])
            n_sl = u64(); slices = np.frombuffer(take(16 * n_sl), "<u8").reshape(n_sl, 2)
            n_col = u64(); column_ids = np.frombuffer(take(n_col), np.uint8)
            if not (n == n_ent == n_sl == n_col):
                raise ValueError("inconsistent lengths")
            rowptr = np.zeros(n + 1, np.int64)
            if n:
                rowptr[:-1], rowptr[-1] = slices[:, 0], slices[-1, 1]
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"Deserialization failed: {e}") from None
        hashes = np.array([_lib.lib().cleora_hash_entity(s.encode("utf-8"), len(s.encode("utf-8"))) for s in ids],
                          np.uint64)
        new = SparseMatrix.from_csr(rowptr, edges["c"], edges["l"], edges["s"], row_sum, hashes)
        old, self._h = getattr(self, "_h", None), new._h
        new._h = None
        if old:
            _lib.lib().cleora_graph_destroy(old)
        self._csr_cache = None
        self._ids_cache = None
        self.entity_ids = ids
        L = _lib.lib()
        check(L.cleora_graph_set_descriptor(self._h, col_a_id, name_a, col_b_id, name_b))
        cids = np.ascontiguousarray(column_ids)
        check(L.cleora_graph_set_column_ids(self._h, ptr(cids, _lib.c_u8p), n))
