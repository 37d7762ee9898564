"""ctypes prototypes for libcleora_b200.so (include/cleora_b200.h).

This module is the *only* place the shared library is loaded.  It fails loudly when the library has not been
built -- there is no Python/numpy fallback for any compute entry point (the CPU oracle under ``oracle/`` is test
infrastructure and is never imported from here).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcleora_b200.so")

OK, ERR_VALUE, ERR_RUNTIME, ERR_CUDA = 0, 1, 2, 3
MARKOV = {"left": 0, "symmetric": 1}
NORM_NONE, NORM_L2_RUST, NORM_L2_NUMPY, NORM_L1_NUMPY = 0, 1, 2, 3

c_i64p = C.POINTER(C.c_int64)
c_u32p = C.POINTER(C.c_uint32)
c_u64p = C.POINTER(C.c_uint64)
c_u8p = C.POINTER(C.c_uint8)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)

EIGH_FN = C.CFUNCTYPE(C.c_int, c_f64p, c_f64p, C.c_int64, C.c_void_p)

# name -> (restype, argtypes); every symbol declared in include/cleora_b200.h
PROTOTYPES = {
    "cleora_last_error": (C.c_char_p, []),
    "cleora_version": (C.c_char_p, []),
    "cleora_device_count": (C.c_int, []),
    "cleora_set_device": (C.c_int, [C.c_int]),
    "cleora_graph_from_lines": (C.c_int, [C.c_char_p, c_i64p, C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "cleora_graph_from_files": (C.c_int, [C.POINTER(C.c_char_p), C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "cleora_graph_from_pairs": (C.c_int, [c_u32p, c_u32p, C.c_int64, C.c_char_p, C.POINTER(C.c_void_p)]),
    "cleora_graph_from_hyperedges": (C.c_int, [c_u32p, c_i64p, C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "cleora_dev_graph_from_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.POINTER(C.c_void_p), c_i64p]),
    "cleora_dev_synth_pairs": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_uint64, C.c_double, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "cleora_dev_graph_hashes": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), c_i64p]),
    "cleora_graph_num_entities_global": (C.c_int64, [C.c_void_p]),
    "cleora_graph_from_csr": (C.c_int, [c_i64p, c_u32p, c_f32p, c_f32p, c_f32p, c_u64p, C.c_int64, C.c_int64,
                                        C.c_int64, C.POINTER(C.c_void_p)]),
    "cleora_graph_destroy": (None, [C.c_void_p]),
    "cleora_graph_release_device": (C.c_int, [C.c_void_p]),
    "cleora_graph_refresh_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cleora_graph_num_entities": (C.c_int64, [C.c_void_p]),
    "cleora_graph_num_cols": (C.c_int64, [C.c_void_p]),
    "cleora_graph_num_edges": (C.c_int64, [C.c_void_p]),
    "cleora_graph_copy_csr": (C.c_int, [C.c_void_p, c_i64p, c_u32p, c_f32p, c_f32p]),
    "cleora_graph_copy_row_sums": (C.c_int, [C.c_void_p, c_f32p]),
    "cleora_graph_copy_entity_hashes": (C.c_int, [C.c_void_p, c_u64p]),
    "cleora_graph_copy_column_ids": (C.c_int, [C.c_void_p, c_u8p]),
    "cleora_graph_entity_ids_nbytes": (C.c_int64, [C.c_void_p]),
    "cleora_graph_copy_entity_ids": (C.c_int, [C.c_void_p, C.c_char_p, c_i64p]),
    "cleora_graph_set_entity_ids": (C.c_int, [C.c_void_p, C.c_char_p, c_i64p, C.c_int64]),
    "cleora_graph_set_descriptor": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p]),
    "cleora_graph_set_column_ids": (C.c_int, [C.c_void_p, c_u8p, C.c_int64]),
    "cleora_graph_col_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "cleora_graph_col_id": (C.c_int, [C.c_void_p, C.c_int]),
    "cleora_graph_find_entity": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int64]),
    "cleora_hash_entity": (C.c_uint64, [C.c_char_p, C.c_int64]),
    "cleora_initialize_deterministically": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_f32p]),
    "cleora_markov_propagate": (C.c_int, [C.c_void_p, c_f32p, C.c_int64, C.c_int64, C.c_int, c_f32p]),
    "cleora_l2_normalize": (C.c_int, [c_f32p, C.c_int64, C.c_int64, c_f32p]),
    "cleora_embed_fast": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_float, c_f32p]),
    "cleora_embed_fast_convergence": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_float,
                                                C.c_float, c_f32p, c_i64p]),
    "cleora_whiten_embeddings": (C.c_int, [c_f32p, C.c_int64, C.c_int64, C.c_int64, c_f32p]),
    "cleora_spectral_rotate": (C.c_int, [c_f32p, C.c_int64, C.c_int64, c_f32p]),
    "cleora_embed": (C.c_int, [C.c_void_p, c_f32p, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_double, C.c_double,
                               C.c_int, C.c_int, c_f32p, c_i64p, c_f64p]),
    "cleora_embed_multi": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int, c_f32p, C.c_int64, C.c_int64, C.c_int,
                                     C.c_int64, C.c_double, C.c_double, C.c_int, C.c_int, c_f32p, c_i64p]),
    "cleora_embed_multi_supported": (C.c_int, [C.c_int64, C.c_int]),
    "cleora_set_eigh": (None, [EIGH_FN, C.c_void_p]),
    "cleora_set_eigh_thread": (None, [C.c_int, EIGH_FN, C.c_void_p]),
    "cleora_set_option": (C.c_int, [C.c_char_p, C.c_int64]),
    "cleora_get_option": (C.c_int64, [C.c_char_p]),
    "cleora_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "cleora_host_free": (None, [C.c_void_p]),
    "cleora_dev_graph_prepare": (C.c_int, [C.c_void_p]),
    "cleora_dev_init": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "cleora_dev_spmm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float,
                                  C.c_float, C.c_int, C.c_void_p]),
    "cleora_dev_spmm_push": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p), C.c_int,
                                       C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "cleora_dev_whiten_apply_push": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                               C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cleora_dev_spmm_scatter": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_int, C.c_int64,
                                          C.c_int64, C.c_int64, C.c_void_p, C.c_float, C.c_float, C.c_void_p]),
    "cleora_dev_whiten_apply_slices": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                                 C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int, C.c_void_p,
                                                 C.c_int, C.c_void_p]),
    "cleora_dev_normalize_slices": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.POINTER(C.c_void_p),
                                              C.c_int, C.c_int64, C.c_void_p]),
    "cleora_dev_malloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "cleora_dev_free": (C.c_int, [C.c_void_p]),
    "cleora_ipc_get_handle": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cleora_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "cleora_ipc_close": (C.c_int, [C.c_void_p]),
    "cleora_dev_normalize": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "cleora_dev_col_sums": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    "cleora_dev_centered_gram": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cleora_dev_whiten_apply": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_void_p]),
    "cleora_dev_whiten_apply_ex": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                             C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "cleora_dev_row_scale": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "cleora_whiten_apply_fusable": (C.c_int, [C.c_int64, C.c_int64]),
    "cleora_dev_sq_diff_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "cleora_dev_whiten_transform": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "cleora_dev_chol_whiten": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cleora_whiten_transform_from_cov": (C.c_int, [c_f64p, C.c_int64, C.c_int64, c_f32p]),
    "cleora_release_workspace": (C.c_int, []),
    "cleora_dev_workspace_bytes": (C.c_int64, []),
    "cleora_kernel_launch_count": (C.c_int64, []),
}

_lib = None
_eigh_keepalive = None


_blas_ctl = None


def _numpy_eigh(a_ptr, w_ptr, d, _user):
    """numpy.linalg.eigh behind the C callback -- the very call the reference makes
    (pycleora/__init__.py:145), so eigenvector signs/order follow the same LAPACK.  The d x d problem is far too
    small for a 100+-thread BLAS pool (measured at d=256 on the 128-thread GPU host: 17-55 ms unrestricted, 5.1 ms
    with 4 threads, 4.6 ms with 1; profiles/r1h_eigh_probe.txt), so the call runs under a thread limit
    (CLEORA_B200_EIGH_THREADS, default 1)."""
    global _blas_ctl
    try:
        a = np.ctypeslib.as_array(a_ptr, shape=(d, d))
        w = np.ctypeslib.as_array(w_ptr, shape=(d,))
        if _blas_ctl is None:
            try:
                from threadpoolctl import ThreadpoolController
                _blas_ctl = ThreadpoolController()
            except Exception:  # noqa: BLE001
                _blas_ctl = False
        if _blas_ctl:
            with _blas_ctl.limit(limits=int(os.environ.get("CLEORA_B200_EIGH_THREADS", "1")), user_api="blas"):
                vals, vecs = np.linalg.eigh(a)
        else:
            vals, vecs = np.linalg.eigh(a)
        a[:, :] = vecs
        w[:] = vals
        return 0
    except Exception:  # noqa: BLE001 - must not propagate through the C frame
        return 1


def lib():
    """Load libcleora_b200.so (once).  Raises ImportError with build instructions when it is missing."""
    global _lib, _eigh_keepalive
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C cleora_b200/csrc`.  cleora_b200 has no CPU/numpy fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(L, name)       # AttributeError here means header and library are out of sync
        fn.restype = res
        fn.argtypes = args
    _eigh_keepalive = EIGH_FN(_numpy_eigh)
    if eigh_mode() == "numpy":
        L.cleora_set_eigh(_eigh_keepalive, None)
    if os.environ.get("CLEORA_B200_PIPELINE", "1") == "0":
        L.cleora_set_option(b"pipeline_whiten", 0)
    if os.environ.get("CLEORA_B200_CHOL", "1") == "0":
        L.cleora_set_option(b"chol_whiten", 0)
    _lib = L
    return L


def eigh_mode() -> str:
    """CLEORA_B200_EIGH: "auto" (default), "numpy" (always the reference's LAPACK call, on the host) or "cusolver"
    (always cuSOLVER Dsyevd on the device).  The loop itself no longer needs an eigensolver per iteration (iterates
    that stay inside the loop are whitened with the Cholesky factor on the device, see chol_whiten.cu); the choice
    matters for the iterates that leave it.  auto = cuSOLVER, except where the reference's LAPACK sign convention
    decides the outcome (rmse early stop on whitened iterates compares successive iterates element-wise)."""
    m = os.environ.get("CLEORA_B200_EIGH", "auto")
    return m if m in ("auto", "numpy", "cusolver") else "auto"


def auto_host_eigh(n: int, d: int, iters: int, norm: int, whiten: bool, residual_weight: float,
                   convergence_threshold: float) -> bool:
    """Does this call need LAPACK's conventions on the host (see eigh_mode)?"""
    if not whiten or eigh_mode() == "cusolver":
        return False
    return convergence_threshold > 0


class host_eigh:
    """Context manager: route the whitening eigensolve through numpy's LAPACK (the reference's call) while active."""

    def __init__(self, enable: bool = True):
        self.enable = enable and eigh_mode() != "numpy"

    def __enter__(self):
        if self.enable:
            lib().cleora_set_eigh_thread(1, _eigh_keepalive, None)
        return self

    def __exit__(self, *exc):
        if self.enable:
            lib().cleora_set_eigh_thread(0, C.cast(None, EIGH_FN), None)
        return False


def check(rc: int) -> None:
    """Map a status code to the exception class the reference raises."""
    if rc == OK:
        return
    msg = lib().cleora_last_error().decode("utf-8", "replace")
    if rc == ERR_VALUE:
        raise ValueError(msg)
    raise RuntimeError(msg)


def ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(ctype)


def f32p(address: int):
    """A raw address (e.g. torch.Tensor.data_ptr()) as float*."""
    return C.cast(C.c_void_p(address), c_f32p)
