"""cleora_b200 -- B200-native drop-in for pycleora's iterated Markov-propagation embedding loop.

Host-side mirror of the reference's public hot-path API (``pycleora/__init__.py``): ``SparseMatrix``,
``embed()``, ``whiten_embeddings()``, ``embed_using_baseline_cleora()`` with identical signatures and error
behaviour.  Every numerical step runs in hand-written sm_100a CUDA kernels behind the C ABI in
``include/cleora_b200.h``; nothing here falls back to numpy for compute (the only numpy call on the path is the
``d x d`` ``numpy.linalg.eigh`` that the reference itself makes, served to the library through a callback).
"""
from __future__ import annotations

from typing import Callable, Optional, Union

import numpy as np

from . import _lib
from ._lib import check, ptr
from .pycleora import SparseMatrix, get_devices, set_devices

__all__ = ["SparseMatrix", "embed", "whiten_embeddings", "embed_using_baseline_cleora", "embed_multiscale",
           "embed_with_node_features", "embed_inductive", "update_graph", "pinned_empty", "set_option", "synth_pairs",
           "release_workspace", "set_devices", "get_devices", "DEFAULT_FEATURE_DIM", "DEFAULT_NUM_ITERATIONS"]

DEFAULT_FEATURE_DIM = 256          # pycleora/__init__.py:12
DEFAULT_NUM_ITERATIONS = 40        # pycleora/__init__.py:13

_DEVICE_NORMS = {"l2": _lib.NORM_L2_NUMPY, "l1": _lib.NORM_L1_NUMPY, "none": _lib.NORM_NONE}


def set_option(key: str, value: int) -> None:
    """Library tuning switches (include/cleora_b200.h: cleora_set_option), e.g. ``set_option("pipeline_whiten", 0)``
    keeps the reference's stage order exactly instead of overlapping the eigensolve with the next SpMM."""
    check(_lib.lib().cleora_set_option(key.encode(), int(value)))


def release_workspace() -> None:
    """Free the calling thread's cached device buffers (iterates, whitening scratch, cuSOLVER workspace)."""
    check(_lib.lib().cleora_release_workspace())


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """A numpy array over page-locked host memory (cudaMallocHost) for fast result downloads."""
    import ctypes as C
    import weakref
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(_lib.lib().cleora_host_alloc(max(nbytes, 1), C.byref(p)))
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, _lib.lib().cleora_host_free, p)
    return arr


def synth_pairs(kind: str, n_nodes: int, n_pairs: int, seed: int = 0, alpha: float = 0.5):
    """Synthetic edge list generated on the current GPU (counter-based RNG, include/cleora_b200.h:
    cleora_dev_synth_pairs): ``kind`` "er" (uniform endpoints) or "chunglu" (endpoint weights (i+10)^-alpha).
    Returns two torch int32 CUDA tensors holding the unsigned 32-bit ids; u != v."""
    import torch
    kinds = {"er": 0, "chunglu": 1}
    if kind not in kinds:
        raise ValueError(f"unknown synthetic graph kind '{kind}'")
    u = torch.empty(int(n_pairs), dtype=torch.int32, device="cuda")
    v = torch.empty(int(n_pairs), dtype=torch.int32, device="cuda")
    check(_lib.lib().cleora_dev_synth_pairs(kinds[kind], int(n_nodes), int(n_pairs), int(seed), float(alpha),
                                            u.data_ptr(), v.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return u, v


def _validate_propagation(propagation: str) -> None:
    if propagation not in ("left", "symmetric"):   # pycleora/__init__.py:24-26
        raise ValueError(f"Unknown propagation type: '{propagation}'. Use 'left' or 'symmetric'.")


def whiten_embeddings(embeddings: np.ndarray, n_components: Optional[int] = None) -> np.ndarray:
    """PCA whitening, ``pycleora/__init__.py:130-164``: f64 mean and covariance (K2), ``eigh``, f32 transform (K3)."""
    x = np.ascontiguousarray(embeddings, dtype=np.float32)
    n, d = x.shape
    if n <= 1:
        return embeddings.copy()
    # `eigenvectors[:, :n_components]` (pycleora/__init__.py:151-153): slice semantics, including 0 and negatives
    dout = d if n_components is None else len(range(d)[:int(n_components)])
    out = np.empty((n, dout), np.float32)
    if dout == 0:
        return out
    check(_lib.lib().cleora_whiten_embeddings(ptr(x, _lib.c_f32p), n, d, dout, ptr(out, _lib.c_f32p)))
    return out


def embed_using_baseline_cleora(graph: SparseMatrix, feature_dim: int, iter: int) -> np.ndarray:  # noqa: A002
    """pycleora/__init__.py:16-21."""
    return embed(graph, feature_dim, iter, whiten=True)


def embed(
    graph: SparseMatrix,
    feature_dim: int = DEFAULT_FEATURE_DIM,
    num_iterations: Union[int, str] = DEFAULT_NUM_ITERATIONS,
    propagation: str = "left",
    normalization: str = "l2",
    seed: int = 0,
    initial_embeddings: Optional[np.ndarray] = None,
    num_workers: Optional[int] = None,
    callback: Optional[Callable[[int, np.ndarray], None]] = None,
    residual_weight: float = 0.0,
    convergence_threshold: float = 0.0,
    whiten: bool = True,
) -> np.ndarray:
    """``pycleora.embed`` (pycleora/__init__.py:51-127) with the whole loop resident on the GPU.

    Dispatch (results as the reference computes them):
      * ``whiten=False``, l2, no callback / initial embeddings -> ``graph.embed_fast[_convergence]`` (Rust fast path);
      * no callback -> one device-resident call for all iterations (``cleora_embed``);
      * callback given -> the reference's per-iteration loop, each stage still on the device, with the host array
        handed to the callback after every iteration.
    """
    if isinstance(num_iterations, str):
        if num_iterations == "auto":
            num_iterations = DEFAULT_NUM_ITERATIONS
        else:
            raise ValueError(f"num_iterations must be an int or 'auto', got '{num_iterations}'")
    use_fast_path = initial_embeddings is None and callback is None and normalization == "l2" and not whiten
    if use_fast_path:
        if convergence_threshold > 0:
            embeddings, _ = graph.embed_fast_convergence(
                feature_dim, num_iterations, propagation=propagation, seed=seed, residual_weight=residual_weight,
                convergence_threshold=convergence_threshold, num_workers=num_workers)
            return embeddings
        return graph.embed_fast(feature_dim, num_iterations, propagation=propagation, seed=seed,
                                residual_weight=residual_weight, num_workers=num_workers)

    _validate_propagation(propagation)
    if normalization not in ("l2", "l1", "none", "spectral"):
        raise ValueError(f"Unknown normalization method: {normalization}. Use 'l2', 'l1', 'spectral', or 'none'.")
    # normalization="spectral" (pycleora/__init__.py:951-956) = row l2 norm followed by a rotation of the iterate onto
    # its right singular vectors.  The loop body is equivariant under orthogonal right-multiplication, so the loop runs
    # with plain l2 and the rotation is applied to the iterates that leave it; with whiten=True the PCA whitening absorbs
    # it altogether (checked against the reference's output in tests/test_gpu_parity.py).
    spectral = normalization == "spectral"
    if spectral:
        normalization = "l2"

    def _rotate(e):
        if not spectral or whiten:
            return e
        out = np.empty_like(e)
        check(_lib.lib().cleora_spectral_rotate(ptr(np.ascontiguousarray(e), _lib.c_f32p), e.shape[0], e.shape[1],
                                                ptr(out, _lib.c_f32p)))
        return out
    if initial_embeddings is not None:
        x0 = initial_embeddings.astype(np.float32)
        if x0.shape[0] != graph.num_entities:
            raise ValueError(
                f"initial_embeddings has {x0.shape[0]} rows but graph has {graph.num_entities} entities")
    else:
        x0 = None

    # Eigensolver choice (see _lib.eigh_mode): the rmse early stop on the whitened path compares successive iterates
    # element-wise, so it depends on the eigensolver's sign conventions -- such calls use the reference's own LAPACK
    # eigh; the pipelined default loop also uses it (on the host, hidden behind the SpMM); otherwise cuSOLVER.
    lapack = whiten and convergence_threshold > 0 and _lib.eigh_mode() != "cusolver"
    if callback is None and not (spectral and not whiten and convergence_threshold > 0):
        out, _ = graph.embed_device(feature_dim, num_iterations, propagation, _DEVICE_NORMS[normalization], seed,
                                    x0, residual_weight, convergence_threshold, whiten)
        return _rotate(out) if num_iterations > 0 else out

    # per-iteration path: one device-resident iteration at a time so the callback sees every iterate
    embeddings = x0 if x0 is not None else graph.initialize_deterministically(feature_dim, seed)
    for i in range(num_iterations):
        prev = embeddings
        with _lib.host_eigh(lapack):
            embeddings, _ = graph.embed_device(embeddings.shape[1], 1, propagation, _DEVICE_NORMS[normalization],
                                               seed, embeddings, residual_weight, 0.0, whiten)
            embeddings = _rotate(embeddings)
        if callback is not None:
            callback(i, embeddings)
        if convergence_threshold > 0 and i > 0:
            diff = embeddings.astype(np.float64, copy=False) - prev.astype(np.float64, copy=False)
            if float(np.sqrt(np.mean(diff * diff))) < convergence_threshold:   # _compute_rmse, :974-976
                break
    return embeddings


from .callers import embed_inductive, embed_multiscale, embed_with_node_features, update_graph  # noqa: E402
