"""In-process multi-GPU embed() (cleora_embed_multi, cleora_b200/csrc/multi_gpu.inl) on the GPU box (`-m gpu`).

The device list may name one GPU several times -- several ranks (host threads) then share it, which exercises the whole
choreography (column slices, peer stores into other ranks' buffers, event barriers, peer-sum all-reduce, redundant
Cholesky factors, PCA on rank 0 + copy) on the driver's single-GPU box.  With >= 2 visible GPUs the same cases also run on
distinct devices over NVLink peer access.
Bars (same as tests/test_gpu_sharded.py): whiten=False is bit-identical to the single-GPU path for every rank count; the
whitened loop agrees with the single-GPU path and with the CPU oracle in Procrustes (<= 1e-4) and Gram (<= 1e-5) terms."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    dict(feature_dim=64, num_iterations=10, whiten=False),                                  # Rust fast path semantics
    dict(feature_dim=128, num_iterations=6, whiten=False, residual_weight=0.2, propagation="symmetric"),
    dict(feature_dim=64, num_iterations=12, whiten=False, convergence_threshold=1e-3),      # f32 rmse early stop
    dict(feature_dim=64, num_iterations=5, whiten=False, normalization="l1"),               # Python loop, no whitening
    dict(feature_dim=64, num_iterations=8, whiten=True),                                    # pipelined, Cholesky inner iterations
    dict(feature_dim=256, num_iterations=6, whiten=True),
    dict(feature_dim=128, num_iterations=5, whiten=True, residual_weight=0.3),              # reference stage order
    dict(feature_dim=64, num_iterations=4, whiten=True, normalization="none"),
]


@pytest.fixture(scope="module")
def world():
    import cleora_b200 as cb
    import oracle
    from tests.helpers import er_lines
    lines, columns = er_lines(30000, 400000, 5), "complex::reflexive::node"
    g = cb.SparseMatrix.from_iterator(lines, columns)
    o = oracle.build_graph(lines, columns)
    cb.set_devices(None)
    single = [cb.embed(g, **kw) for kw in CASES]
    refs = [oracle.embed(o, **kw) if kw["whiten"] else None for kw in CASES]
    yield g, single, refs
    cb.set_devices(None)


def _check(g, single, refs, devices):
    import cleora_b200 as cb
    from tests.helpers import gram_err, procrustes_err
    cb.set_devices(devices)
    try:
        for kw, one, ref in zip(CASES, single, refs):
            out = cb.embed(g, **kw)
            assert out.shape == one.shape, kw
            if not kw["whiten"]:
                np.testing.assert_array_equal(out, one, err_msg=f"{kw} on devices {devices}")
            else:
                for other, name in ((one, "single GPU"), (ref, "oracle")):
                    assert procrustes_err(out, other) <= 1e-4, (kw, name, devices)
                    assert gram_err(out, other) <= 1e-5, (kw, name, devices)
    finally:
        cb.set_devices(None)


@pytest.mark.parametrize("ranks", [2, 4, 8])
def test_ranks_sharing_one_gpu(world, ranks):
    g, single, refs = world
    _check(g, single, refs, [0] * ranks)


def test_all_visible_gpus(world):
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus N)")
    g, single, refs = world
    _check(g, single, refs, list(range(n)))


def test_initial_embeddings_early_stop_and_iteration_count(world):
    """x0 given (column slices are cut from the host matrix), the whitened loop with the rmse check armed (PCA basis in
    every iteration, LAPACK sign conventions through the host callback on rank 0; the threshold is never met), the
    fast path's early stop, and iters_done."""
    import cleora_b200 as cb
    g, _, _ = world
    x0 = np.random.default_rng(3).standard_normal((g.num_entities, 64)).astype(np.float32)
    cb.set_devices(None)
    one, it_one = g.embed_device(64, 8, "left", initial_embeddings=x0, convergence_threshold=1e-9, whiten=True)
    plain_one, _ = g.embed_device(64, 3, "left", initial_embeddings=x0, whiten=False)
    cb.set_devices([0, 0, 0, 0])
    try:
        out, it_multi = g.embed_device(64, 8, "left", initial_embeddings=x0, convergence_threshold=1e-9, whiten=True)
        plain, _ = g.embed_device(64, 3, "left", initial_embeddings=x0, whiten=False)
        zero, it_zero = g.embed_device(64, 0, "left", initial_embeddings=x0, whiten=True)
    finally:
        cb.set_devices(None)
    np.testing.assert_array_equal(plain, plain_one)
    cb.set_devices(None)
    fast_one, n_one = g.embed_fast_convergence(64, 40, convergence_threshold=3e-3)
    cb.set_devices([0, 0])
    try:
        fast, n_multi = g.embed_fast_convergence(64, 40, convergence_threshold=3e-3)
    finally:
        cb.set_devices(None)
    assert n_multi == n_one
    np.testing.assert_array_equal(fast, fast_one)
    np.testing.assert_array_equal(zero, x0)
    assert it_zero == 0
    assert it_multi == it_one == 8
    from tests.helpers import procrustes_err
    assert procrustes_err(out, one) <= 1e-4


def test_unsupported_shapes_and_devices(world):
    import cleora_b200 as cb
    from cleora_b200 import _lib
    g, _, _ = world
    L = _lib.lib()
    assert L.cleora_embed_multi_supported(256, 8) == 1 and L.cleora_embed_multi_supported(48, 2) == 0
    cb.set_devices([0, 0])
    try:
        a = cb.embed(g, 48, 3, whiten=False)              # 24-float slices: no slice kernel -> one-GPU path, same result
    finally:
        cb.set_devices(None)
    np.testing.assert_array_equal(a, cb.embed(g, 48, 3, whiten=False))
    import ctypes as C
    out = np.empty((g.num_entities, 64), np.float32)
    devs = (C.c_int * 2)(0, 99)
    rc = L.cleora_embed_multi(g._handle(), devs, 2, None, 64, 2, 0, 0, 0.0, 0.0, _lib.NORM_L2_NUMPY, 1,
                              _lib.ptr(out, _lib.c_f32p), None)
    assert rc == _lib.ERR_VALUE and b"not visible" in L.cleora_last_error()
    rc = L.cleora_embed_multi(g._handle(), devs, 2, None, 48, 2, 0, 0, 0.0, 0.0, _lib.NORM_L2_NUMPY, 1,
                              _lib.ptr(out, _lib.c_f32p), None)
    assert rc == _lib.ERR_VALUE
