"""Parity of the CUDA path against the CPU oracle and the committed reference goldens.
Everything here calls through the C ABI (via the ctypes SparseMatrix mirror) and needs a B200: `-m gpu`.

Tolerances (stated per test):
  * integer / index / init / un-normalised SpMM: bit-exact;
  * L2-normalised iterates, whiten=False, 40 iterations: <= 1e-5 relative per element (elements above 1% of the
    matrix rms) and <= 1e-5 of the matrix scale everywhere;
  * whitening stages teacher-forced: mean/cov <= 1e-12 relative (f64), apply-T <= 1e-5 of scale;
  * whitened loop end to end: raw per-element where the spectrum is well separated, otherwise after orthogonal
    Procrustes alignment and through the Gram matrix (the PCA eigenbasis is ill-conditioned, SURVEY.md finding 3).
"""
import ast
import os

import numpy as np
import pytest

import cleora_b200 as cb
import oracle
from cleora_b200 import _lib
from tests.helpers import KARATE_COLUMNS, KARATE_EDGES, er_lines, gram_err, procrustes_err, scale_rel_err

pytestmark = pytest.mark.gpu


def _pair(lines, columns, trim=16):
    return cb.SparseMatrix.from_iterator(lines, columns, hyperedge_trim_n=trim), oracle.build_graph(lines, columns, trim)


@pytest.fixture(scope="module")
def er_pair():
    return _pair(er_lines(20000, 200000, 7), "complex::reflexive::node")


@pytest.fixture(scope="module")
def skew_pair():
    """Power-law-ish hypergraph with a few very long rows (long-row path) and many short ones."""
    rs = np.random.default_rng(17)
    w = 1.0 / np.arange(1, 6001) ** 0.9
    w /= w.sum()
    lines = []
    for _ in range(30000):
        k = int(rs.integers(2, 9))
        lines.append(" ".join(str(int(t)) for t in rs.choice(6000, size=k, p=w)))
    return _pair(lines, "complex::reflexive::n")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ------------------------------------------------------------------------------------------------ P1: init
@pytest.mark.parametrize("d,seed", [(1, 0), (7, 3), (32, 0), (256, -5), (300, 2**40 + 1)])
def test_init_bit_exact(er_pair, d, seed):
    g, o = er_pair
    np.testing.assert_array_equal(bits(g.initialize_deterministically(d, seed)), bits(oracle.init_matrix(o.hashes, d, seed)))


# ------------------------------------------------------------------------------------------------ P2: snapshots
@pytest.mark.parametrize("tag,markov,key", [("01", "left", "left_01"), ("02", "left", "left_02"),
                                            ("01", "symmetric", "sym_01"), ("02", "symmetric", "sym_02")])
def test_reference_snapshots_on_gpu(golden_dir, tag, markov, key):
    """tests/snapshot.rs:18-50 through the CUDA SpMM: all 3200 integers of each insta snapshot."""
    z = np.load(os.path.join(golden_dir, "snapshots.npz"))
    g = cb.SparseMatrix.from_iterator([str(s) for s in z[f"lines_{tag}"]], str(z[f"columns_{tag}"]))
    fn = g.left_markov_propagate if markov == "left" else g.symmetric_markov_propagate
    got = (fn(z[f"emb_{tag}"]) * np.float32(1000.0)).astype(np.int32)
    np.testing.assert_array_equal(got, z[key])


# ------------------------------------------------------------------------------------------------ SpMM bit-exact
@pytest.mark.parametrize("d", [1, 3, 8, 16, 32, 64, 96, 100, 128, 192, 256, 260, 384, 512, 1024])
@pytest.mark.parametrize("markov", ["left", "symmetric"])
def test_spmm_bit_exact_all_widths(er_pair, d, markov):
    g, o = er_pair
    x = np.random.default_rng(d).standard_normal((o.n, d)).astype(np.float32)
    fn = g.left_markov_propagate if markov == "left" else g.symmetric_markov_propagate
    np.testing.assert_array_equal(bits(fn(x)), bits(oracle.spmm(o, x, markov)))


@pytest.mark.parametrize("d", [32, 128, 256])
def test_spmm_bit_exact_skewed_rows(skew_pair, d):
    g, o = skew_pair
    assert np.diff(o.rowptr).max() > 2000
    x = np.random.default_rng(1).standard_normal((o.n, d)).astype(np.float32)
    np.testing.assert_array_equal(bits(g.left_markov_propagate(x)), bits(oracle.spmm(o, x, "left")))


def test_hub_rows_are_split_deterministically():
    """Degree skew (SURVEY.md hard part): rows above the long-row threshold (default 65536 edges; lowered to 8192
    here) are summed chunk-wise by several warps -- deterministic, every other row still bit-exact.  The chunked f32
    sum differs from the reference's sequential sum by ~sqrt(deg) ulp on those rows, which is why the default
    threshold is high: with the default, this graph stays bit-exact and the 40-iteration loop stays within 1e-5."""
    rs = np.random.default_rng(3)
    n_leaf = 30000
    lines = [f"hub {i}" for i in range(n_leaf)] + [f"hub2 {i}" for i in range(0, n_leaf, 3)]
    lines += [f"{int(a)} {int(b)}" for a, b in rs.integers(0, n_leaf, size=(60000, 2))]
    o = oracle.build_graph(lines, "complex::reflexive::n")
    deg = np.diff(o.rowptr)
    assert deg.max() > 8192 and (deg > 8192).sum() == 2
    x = rs.standard_normal((o.n, 256)).astype(np.float32)
    ref = oracle.spmm(o, x)
    g = cb.SparseMatrix.from_iterator(lines, "complex::reflexive::n")                  # default threshold: no split
    np.testing.assert_array_equal(bits(g.left_markov_propagate(x)), bits(ref))
    _assert_1e5(g.embed_fast(128, 40), oracle.embed_fast(o, 128, 40))
    os.environ["CLEORA_B200_LONG_ROW"], os.environ["CLEORA_B200_LONG_CHUNK"] = "8192", "2048"
    try:
        gs = cb.SparseMatrix.from_iterator(lines, "complex::reflexive::n")             # schedule is built at upload
        got = gs.left_markov_propagate(x)
    finally:
        del os.environ["CLEORA_B200_LONG_ROW"], os.environ["CLEORA_B200_LONG_CHUNK"]
    short = deg <= 8192
    np.testing.assert_array_equal(bits(got[short]), bits(ref[short]))
    np.testing.assert_allclose(got[~short], ref[~short], rtol=1e-4, atol=1e-6 * np.abs(ref).max())
    np.testing.assert_array_equal(bits(gs.left_markov_propagate(x)), bits(got))        # run-to-run identical
    x128 = np.ascontiguousarray(x[:, :128])
    got128 = gs.left_markov_propagate(x128)
    np.testing.assert_array_equal(bits(got128[short]), bits(oracle.spmm(o, x128)[short]))
    x32 = np.ascontiguousarray(x[:, :32])                              # narrow rows (several rows per warp) split the same way
    got32, ref32 = gs.left_markov_propagate(x32), oracle.spmm(o, x32)
    np.testing.assert_array_equal(bits(got32[short]), bits(ref32[short]))
    np.testing.assert_allclose(got32[~short], ref32[~short], rtol=1e-4, atol=1e-6 * np.abs(ref32).max())


def test_push_epilogues_replicate_rows(er_pair):
    """The fused-gather hooks (cleora_dev_spmm_push / cleora_dev_whiten_apply_push) store every produced row into the
    extra destinations as well -- exercised here with extra buffers on the same GPU (across GPUs the destinations are
    CUDA-IPC mappings of the peers' buffers: tools/sharded_check.py)."""
    import ctypes as C
    import torch
    g, o = er_pair
    L = _lib.lib()
    n, d = o.n, 256
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((n, d)).astype(np.float32)).cuda()
    outs = [torch.zeros(n, d, device="cuda") for _ in range(4)]
    extra = (C.c_void_p * 3)(*[t.data_ptr() for t in outs[1:]])
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.cleora_dev_graph_prepare(g._handle()))
    _lib.check(L.cleora_dev_spmm_push(g._handle(), 0, x.data_ptr(), d, outs[0].data_ptr(), extra, 3, None, 1.0, 0.0,
                                      _lib.NORM_L2_RUST, st))
    torch.cuda.synchronize()
    ref = oracle.l2_normalize(oracle.spmm(o, x.cpu().numpy()))
    np.testing.assert_allclose(outs[0].cpu().numpy(), ref, rtol=1e-6, atol=1e-9)
    for t in outs[1:]:
        assert torch.equal(t, outs[0])
    # tensor-core GEMM epilogue
    mean = torch.zeros(d, device="cuda")
    T = torch.eye(d, device="cuda").contiguous()
    zs = [torch.zeros(n, d, device="cuda") for _ in range(3)]
    extra2 = (C.c_void_p * 2)(*[t.data_ptr() for t in zs[1:]])
    _lib.check(L.cleora_dev_whiten_apply_push(outs[0].data_ptr(), n, d, mean.data_ptr(), T.data_ptr(), d,
                                              zs[0].data_ptr(), extra2, 2, _lib.NORM_L2_NUMPY, None, st))
    torch.cuda.synchronize()
    np.testing.assert_allclose(zs[0].cpu().numpy(), ref, rtol=0, atol=2e-6)      # identity transform, unit rows
    assert torch.equal(zs[1], zs[0]) and torch.equal(zs[2], zs[0])


@pytest.mark.parametrize("d,G", [(256, 8), (128, 8), (256, 2), (64, 4), (512, 8)])
def test_fused_transposes_of_the_column_sharded_loop(er_pair, d, G):
    """The three kernels behind cleora_b200/colsharded.py with all G "ranks" played by one process on one GPU (their
    multi-process use over CUDA IPC is tests/test_gpu_sharded.py):
      K1 on a column slice, rows scattered to the row owners  == the single product, bit for bit;
      K1's row normalisation with column slices to the slice owners == the fused-norm single-GPU K1, bit for bit;
      K3 with column slices == K3 (same kernel, other destinations), bit for bit."""
    import ctypes as C
    import torch
    g, o = er_pair
    L = _lib.lib()
    n, ds = o.n, d // G
    block = (n + G - 1) // G
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.cleora_dev_graph_prepare(g._handle()))
    x = torch.from_numpy(np.random.default_rng(d + G).standard_normal((n, d)).astype(np.float32)).cuda()
    ptrs = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])   # noqa: E731
    # ---- B -> A: every rank's K1 on its slice, scattered into the owners' W buffers
    wa = [torch.full((block, d), float("nan"), device="cuda") for _ in range(G)]
    for r in range(G):
        xs = x[:, r * ds:(r + 1) * ds].contiguous()
        _lib.check(L.cleora_dev_spmm_scatter(g._handle(), 0, xs.data_ptr(), ds, ptrs(wa), G, block, d, r * ds, None, 1.0, 0.0, st))
    torch.cuda.synchronize()
    w = torch.cat(wa)[:n]
    ref = torch.empty(n, d, device="cuda")
    _lib.check(L.cleora_dev_spmm(g._handle(), 0, x.data_ptr(), d, ref.data_ptr(), None, 1.0, 0.0, _lib.NORM_NONE, st))
    torch.cuda.synchronize()
    assert torch.equal(w, ref)
    np.testing.assert_array_equal(bits(w.cpu().numpy()), bits(oracle.spmm(o, x.cpu().numpy())))
    # ---- A -> B: row normalisation of each row block, column slices to the slice owners
    for norm in (_lib.NORM_L2_RUST, _lib.NORM_L2_NUMPY):
        xb = [torch.full((block * G, ds), float("nan"), device="cuda") for _ in range(G)]
        ya = [torch.empty(block, d, device="cuda") for _ in range(G)]
        for h in range(G):
            rows = min(block, max(0, n - h * block))
            _lib.check(L.cleora_dev_normalize_slices(wa[h].data_ptr(), rows, d, norm, ya[h].data_ptr(), ptrs(xb), G, h * block, st))
        fused = torch.empty(n, d, device="cuda")
        _lib.check(L.cleora_dev_spmm(g._handle(), 0, x.data_ptr(), d, fused.data_ptr(), None, 1.0, 0.0, norm, st))
        torch.cuda.synchronize()
        assert torch.equal(torch.cat(ya)[:n], fused)                         # same summation tree as K1's fused epilogue
        for r in range(G):
            assert torch.equal(xb[r][:n], fused[:, r * ds:(r + 1) * ds])
    # ---- A -> B through the tensor-core apply
    if _lib.lib().cleora_whiten_apply_fusable(d, d):
        rs = np.random.default_rng(1)
        mean = torch.from_numpy((rs.standard_normal(d) * 1e-2).astype(np.float32)).cuda()
        T = torch.from_numpy((rs.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32)).cuda()
        scale = torch.from_numpy(rs.uniform(0.9, 1.1, n).astype(np.float32)).cuda()
        whole = torch.empty(n, d, device="cuda")
        _lib.check(L.cleora_dev_whiten_apply_ex(w.data_ptr(), n, d, mean.data_ptr(), T.data_ptr(), d, whole.data_ptr(),
                                                _lib.NORM_L2_NUMPY, scale.data_ptr(), 0, st))
        xb = [torch.full((block * G, ds), float("nan"), device="cuda") for _ in range(G)]
        ya = [torch.empty(block, d, device="cuda") for _ in range(G)]
        for h in range(G):
            rows = min(block, max(0, n - h * block))
            _lib.check(L.cleora_dev_whiten_apply_slices(wa[h].data_ptr(), rows, d, mean.data_ptr(), T.data_ptr(), d, ya[h].data_ptr(),
                                                        ptrs(xb), G, h * block, _lib.NORM_L2_NUMPY,
                                                        scale[h * block:].data_ptr() if rows else None, 0, st))
        torch.cuda.synchronize()
        assert torch.equal(torch.cat(ya)[:n], whole)
        for r in range(G):
            assert torch.equal(xb[r][:n], whole[:, r * ds:(r + 1) * ds])
        refz = oracle.normalize((w.cpu().numpy() - scale.cpu().numpy()[:, None] * mean.cpu().numpy()) @ T.cpu().numpy(), "l2")
        assert scale_rel_err(whole.cpu().numpy(), refz) <= 1e-5


def test_spmm_edge_cases():
    # single entity, self loop only; empty graph; d = 0
    g, o = _pair(["a"], "complex::reflexive::n")
    x = np.float32([[1.5, -2.0, 0.25]])
    np.testing.assert_array_equal(g.left_markov_propagate(x), oracle.spmm(o, x))
    e = cb.SparseMatrix()
    assert e.left_markov_propagate(np.zeros((0, 4), np.float32)).shape == (0, 4)
    g2, o2 = _pair(KARATE_EDGES, KARATE_COLUMNS)
    assert g2.left_markov_propagate(np.zeros((34, 0), np.float32)).shape == (34, 0)
    # non-contiguous input is copied, not rejected
    xx = np.random.default_rng(0).standard_normal((34, 16)).astype(np.float32)
    np.testing.assert_array_equal(g2.left_markov_propagate(xx[:, ::2]), oracle.spmm(o2, np.ascontiguousarray(xx[:, ::2])))


# ------------------------------------------------------------------------------------------------ L2
@pytest.mark.parametrize("d", [5, 32, 256, 300])
def test_l2_normalize(er_pair, d):
    g, _ = er_pair
    x = np.random.default_rng(d).standard_normal((999, d)).astype(np.float32)
    x[7] = 0.0                                                       # max(norm, 1e-10) guard
    got, ref = g.l2_normalize(x), oracle.l2_normalize(x)
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=0)          # tree vs sequential sum of squares
    np.testing.assert_array_equal(got[7], 0.0)


# ------------------------------------------------------------------------------------------------ P3: fast path
def _assert_1e5(got, ref):
    scale = float(np.sqrt(np.mean(ref.astype(np.float64) ** 2)))
    assert scale_rel_err(got, ref) <= 1e-5
    big = np.abs(ref) >= 1e-2 * scale
    rel = np.abs(got.astype(np.float64) - ref)[big] / np.abs(ref[big])
    assert rel.max() <= 1e-5, rel.max()


@pytest.mark.parametrize("d,markov,rw", [(32, "left", 0.0), (256, "left", 0.0), (128, "symmetric", 0.0),
                                         (64, "left", 0.3), (100, "left", 0.0), (256, "left", 1.5)])
def test_embed_fast_40_iterations_within_1e5(er_pair, d, markov, rw):
    """src/embedding.rs:106-136 x 40: fp32 within 1e-5 relative per element (north_star's bar)."""
    g, o = er_pair
    got = g.embed_fast(d, 40, propagation=markov, seed=3, residual_weight=rw)
    ref = oracle.embed_fast(o, d, 40, markov, 3, rw)
    _assert_1e5(got, ref)
    np.testing.assert_allclose(np.linalg.norm(got.astype(np.float64), axis=1), 1.0, rtol=1e-5)


def test_embed_fast_skewed_and_karate(skew_pair, golden_dir):
    g, o = skew_pair
    _assert_1e5(g.embed_fast(64, 40), oracle.embed_fast(o, 64, 40))
    z = np.load(os.path.join(golden_dir, "embed_karate_d32_t40_now.npz"))
    gk = cb.SparseMatrix.from_iterator([str(s) for s in z["lines"]], str(z["columns"]))
    _assert_1e5(cb.embed(gk, 32, 40, whiten=False), z["out"])       # reference embed() output (whiten=False)


def test_embed_fast_convergence(er_pair):
    g, o = er_pair
    for thr in (0.0, 5e-3, 1e-3):
        got, it = g.embed_fast_convergence(32, 30, convergence_threshold=thr)
        ref, it_ref = oracle.embed_fast_convergence(o, 32, 30, convergence_threshold=thr)
        assert it == it_ref, (thr, it, it_ref)
        _assert_1e5(got, ref)
    assert cb.embed(g, 32, 30, whiten=False, convergence_threshold=5e-3).shape == (o.n, 32)


# ------------------------------------------------------------------------------------------------ P4: whitening stages
def _dev_stats(x):
    """mean (f64) and unscaled centred Gram (f64) of a host matrix through the device-level ABI."""
    import torch
    L = _lib.lib()
    n, d = x.shape
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    sums = torch.zeros(d, dtype=torch.float64, device="cuda")
    cov = torch.zeros(d, d, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.cleora_dev_col_sums(xd.data_ptr(), n, d, sums.data_ptr(), 0, st))
    mean = sums / n
    _lib.check(L.cleora_dev_centered_gram(xd.data_ptr(), n, d, mean.data_ptr(), cov.data_ptr(), st))
    torch.cuda.synchronize()
    return mean.cpu().numpy(), cov.cpu().numpy()


@pytest.mark.parametrize("n,d", [(3000, 48), (4000, 256), (777, 100), (20000, 512), (2, 8), (40, 64)])
def test_mean_and_covariance_f64(n, d):
    rs = np.random.default_rng(n + d)
    x = oracle.normalize((rs.standard_normal((n, d)) * rs.uniform(0.2, 3.0, d) + rs.uniform(-1, 1, d)).astype(np.float32))
    # (d = 512 with n >= 4096 would take the integer tensor-core path since round 2; this test is about the IEEE f64 kernel)
    mean, gram = _with_options(lambda: _dev_stats(x), gram_needed_cols=0) if d > 256 else _dev_stats(x)
    ref_mean, ref_cov = oracle.whiten_stats(x)
    np.testing.assert_allclose(mean, ref_mean, rtol=1e-12, atol=1e-15)
    cov = gram / (n - 1)
    assert np.max(np.abs(cov - ref_cov)) <= 1e-12 * np.max(np.abs(ref_cov))
    np.testing.assert_array_equal(cov, cov.T)


@pytest.fixture(params=["defaults", "k3_asw=1", "k3_asw=0", "k3_bk=16", "gram_needed_cols=1", "gram_needed_cols=0"])
def kernel_variant(request):
    """Runs a test under every selectable kernel variant (cleora_set_option): K3's A-tile layout / stage shape and K2b's
    converted-column set.  Whatever the shipped default is, the alternatives stay green."""
    L = _lib.lib()
    keys = (b"k3_asw", b"k3_bk", b"gram_needed_cols")
    saved = {k: L.cleora_get_option(k) for k in keys}
    if request.param != "defaults":
        k, v = request.param.split("=")
        _lib.check(L.cleora_set_option(k.encode(), int(v)))
    yield request.param
    for k, v in saved.items():
        _lib.check(L.cleora_set_option(k, v))


def test_integer_gram_on_tensor_cores_is_exact(kernel_variant):
    """K2b's tcgen05 kind::i8 path (d in {128, 256}, and 384 / 512 with the compact staging; n >= 4096): for inputs that are exactly representable in its
    fixed-point format the centred Gram matrix must equal the exact rational result -- checked with Python
    integers -- and on generic f32 data it must agree with the f64 oracle to the quantisation bound."""
    rs = np.random.default_rng(5)
    shapes = [(5000, 128), (20000, 256), (4096 + 77, 256)]
    if _lib.lib().cleora_get_option(b"gram_needed_cols") == 1:
        shapes += [(9000, 512), (5000, 384)]                   # four / three row blocks: only with the compact staging
    for n, d in shapes:
        # multiples of 2^-20 in (-1, 1): exactly representable in f32 and in the kernel's 2^-e grid (e >= 29)
        k = rs.integers(-(2 ** 20) + 1, 2 ** 20, size=(n, d))
        k[:, 3] //= 4096                                        # a column of small values (low planes only)
        k[:, 5] = -np.abs(k[:, 5])                              # a sign-constant column (mean far from 0)
        x = (k.astype(np.float64) / 2 ** 20).astype(np.float32)
        assert np.array_equal(x.astype(np.float64) * 2 ** 20, k)
        mean, gram = _dev_stats(x)
        S = k.sum(axis=0)                                        # exact integers
        ref_mean = S.astype(np.float64) / n / 2 ** 20
        np.testing.assert_allclose(mean, ref_mean, rtol=1e-14, atol=1e-18)
        kc = k.astype(np.float64) - S / n                        # exact-ish reference in f64 (|kc| < 2^21, sums < 2^63)
        ref = (kc.T @ kc) / 2.0 ** 40
        assert np.max(np.abs(gram - ref)) <= 1e-13 * np.max(np.abs(ref))
        for i, j in ((3, 5), (0, 0), (5, 5), (d - 1, 7), (3, 3)):  # fully exact check with Python integers
            dot = sum(int(a) * int(b) for a, b in zip(k[:, i], k[:, j]))
            exact = (dot * n - int(S[i]) * int(S[j])) / n / 2 ** 40
            assert abs(gram[i, j] - exact) <= 4e-16 * max(abs(exact), np.max(np.abs(ref)) * 1e-3)
    x = oracle.normalize((rs.standard_normal((30000, 256)) * rs.uniform(0.2, 3.0, 256) + rs.uniform(-1, 1, 256)).astype(np.float32))
    mean, gram = _dev_stats(x)
    ref_mean, ref_cov = oracle.whiten_stats(x)
    cov = gram / (x.shape[0] - 1)
    assert np.max(np.abs(cov - ref_cov)) <= 1e-9 * np.max(np.abs(ref_cov))     # 2^-30 input quantisation
    np.testing.assert_allclose(cov, cov.T, rtol=0, atol=1e-18)


@pytest.mark.parametrize("n,d,dout", [(3000, 48, 48), (5000, 256, 256), (777, 100, 100), (1000, 64, 10), (4097, 512, 512)])
def test_apply_transform_given_identical_T(n, d, dout, kernel_variant):
    import torch
    rs = np.random.default_rng(d)
    x = oracle.normalize(rs.standard_normal((n, d)).astype(np.float32))
    mean, cov = oracle.whiten_stats(x)
    T = oracle.whiten_transform(cov, None if dout == d else dout)
    ref = oracle.whiten_apply(x, mean, T)
    L = _lib.lib()
    xd = torch.from_numpy(x).cuda()
    md = torch.from_numpy(mean.astype(np.float32)).cuda()
    Td = torch.from_numpy(np.ascontiguousarray(T)).cuda()
    out = torch.empty(n, dout, dtype=torch.float32, device="cuda")
    _lib.check(L.cleora_dev_whiten_apply(xd.data_ptr(), n, d, md.data_ptr(), Td.data_ptr(), dout, out.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert scale_rel_err(out.cpu().numpy(), ref) <= 1e-5


@pytest.mark.parametrize("n,d", [(5000, 256), (3000, 128), (4097, 512), (1000, 64)])
def test_apply_with_upper_triangular_transform(n, d, kernel_variant):
    """K3's zero-block skipping for an upper-triangular T (the Cholesky whitening transform) against the dense path of the
    same kernel and against numpy."""
    import torch
    rs = np.random.default_rng(d + 1)
    x = oracle.normalize(rs.standard_normal((n, d)).astype(np.float32))
    mean = (rs.standard_normal(d) * 1e-2).astype(np.float32)
    T = np.triu(rs.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32)
    L = _lib.lib()
    xd, md, Td = (torch.from_numpy(a).cuda() for a in (x, mean, np.ascontiguousarray(T)))
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for upper in (0, 1):
        out = torch.empty(n, d, dtype=torch.float32, device="cuda")
        _lib.check(L.cleora_dev_whiten_apply_ex(xd.data_ptr(), n, d, md.data_ptr(), Td.data_ptr(), d, out.data_ptr(),
                                                _lib.NORM_L2_NUMPY, None, upper, st))
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy())
    ref = oracle.normalize((x - mean) @ T, "l2")
    assert scale_rel_err(outs[0], ref) <= 1e-5 and scale_rel_err(outs[1], ref) <= 1e-5
    assert scale_rel_err(outs[1], outs[0]) <= 2e-7                      # only exact zeros were skipped


def test_whiten_embeddings_matches_reference_golden(golden_dir):
    """Well-separated spectrum (column scales 0.5..3): PCA basis is well conditioned, compare per element."""
    z = np.load(os.path.join(golden_dir, "whiten_stage.npz"))
    got = cb.whiten_embeddings(z["normalized"])
    ref = z["whitened"]
    sign = np.sign(np.sum(got * ref, axis=0))                       # eigenvector sign is a LAPACK convention
    assert scale_rel_err(got * sign, ref) <= 2e-4
    gotk = cb.whiten_embeddings(z["normalized"], n_components=5)
    assert gotk.shape == (3000, 5) and scale_rel_err(gotk * sign[:5], ref[:, :5]) <= 2e-4
    one = np.float32([[1, 2, 3]])
    np.testing.assert_array_equal(cb.whiten_embeddings(one), one)   # n <= 1: copy (pycleora/__init__.py:132-133)


# ------------------------------------------------------------------------------------------------ P5: whitened loop
@pytest.mark.parametrize("name,raw_tol", [("karate_d8_t5_w", 1e-4), ("karate_d32_t5_w", None), ("karate_d8_t40_w", 1e-4)])
def test_default_embed_matches_reference_output_karate(golden_dir, name, raw_tol):
    """pycleora.embed() default path (whiten=True) against the unmodified reference's output."""
    z = np.load(os.path.join(golden_dir, f"embed_{name}.npz"))
    kw = ast.literal_eval(str(z["kwargs"]))
    g = cb.SparseMatrix.from_iterator([str(s) for s in z["lines"]], str(z["columns"]))
    got = cb.embed(g, **kw)
    ref = z["out"]
    sign = np.sign(np.sum(got * ref, axis=0))
    sign[sign == 0] = 1
    if raw_tol is not None:
        assert scale_rel_err(got * sign, ref) <= raw_tol
        assert gram_err(got, ref) <= 1e-4
    # else: karate d=32 (n-1 = 33 ~ d) is chaotic under 1-ulp noise already on the CPU -- the oracle run twice
    # with 6e-8 relative noise differs by O(1) in raw AND Gram terms after 5 iterations -- so
    # only the first iterate is comparable (below).
    # callback path reproduces the same iterates and hands out every one of them
    seen = []
    got_cb = cb.embed(g, callback=lambda i, e: seen.append((i, e.copy())), **kw)
    assert [i for i, _ in seen] == list(range(kw["num_iterations"]))
    np.testing.assert_array_equal(got_cb, seen[-1][1])
    if "trace" in z.files:
        assert gram_err(seen[0][1], z["trace"][0]) <= (1e-5 if raw_tol is not None else 1e-3)


def test_default_embed_er_graph_procrustes_and_gram(golden_dir):
    """Near-degenerate covariance spectrum: raw element-wise agreement is not defined; the
    subspace, the Gram matrix and the Procrustes-aligned iterate are."""
    z = np.load(os.path.join(golden_dir, "embed_er2k_d32_t5_w.npz"))
    g = cb.SparseMatrix.from_iterator([str(s) for s in z["lines"]], str(z["columns"]))
    got, ref = cb.embed(g, 32, 5), z["out"]
    assert procrustes_err(got, ref) <= 1e-4
    assert gram_err(got, ref) <= 1e-5
    c = np.cov(got.astype(np.float64), rowvar=False)
    np.testing.assert_allclose(c, np.eye(32), atol=1e-4)             # whitened: unit covariance
    np.testing.assert_allclose(got.astype(np.float64).mean(0), 0, atol=1e-6)


@pytest.mark.parametrize("name", ["karate_d16_t10_sym_res", "karate_d8_t30_conv", "karate_d16_t6_l1"])
def test_embed_variants_against_reference(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"embed_{name}.npz"))
    kw = ast.literal_eval(str(z["kwargs"]))
    g = cb.SparseMatrix.from_iterator([str(s) for s in z["lines"]], str(z["columns"]))
    got, ref = cb.embed(g, **kw), z["out"]
    assert got.shape == ref.shape                                    # same early-stop iteration => same shape/basis
    assert gram_err(got, ref) <= 1e-3
    assert procrustes_err(got, ref) <= 5e-3


@pytest.mark.parametrize("name", ["karate_d8_t6_spectral_now", "karate_d8_t6_spectral_w"])
def test_spectral_normalization_against_reference(golden_dir, name):
    """normalization="spectral" (pycleora/__init__.py:951-956: row l2 norm, then U*S of the SVD).  The loop is
    equivariant under that rotation, so the device loop runs plain l2 and rotates the iterate that leaves it
    (cleora_spectral_rotate); with whiten=True the PCA whitening absorbs the rotation.  Raw agreement with the
    unmodified reference's output after column-sign alignment."""
    z = np.load(os.path.join(golden_dir, f"embed_{name}.npz"))
    kw = ast.literal_eval(str(z["kwargs"]))
    g = cb.SparseMatrix.from_iterator([str(s) for s in z["lines"]], str(z["columns"]))
    got, ref = cb.embed(g, **kw), z["out"]
    sign = np.sign(np.sum(got * ref, axis=0))
    assert scale_rel_err(got * sign, ref) <= 1e-5
    seen = []
    got_cb = cb.embed(g, callback=lambda i, e: seen.append(e.copy()), **kw)           # per-iteration path
    assert len(seen) == kw["num_iterations"]
    sign = np.sign(np.sum(got_cb * ref, axis=0))
    assert scale_rel_err(got_cb * sign, ref) <= 1e-5


def test_teacher_forced_iteration_on_er_graph(er_pair):
    """One full default iteration from the ORACLE's iterate at t=3: SpMM+L2 <= 1e-6, then whitening compared
    through Gram/Procrustes."""
    g, o = er_pair
    xt = oracle.embed(o, 64, 3)
    y_ref = oracle.normalize(oracle.spmm(o, xt), "l2")
    y, _ = g.embed_device(64, 1, "left", _lib.NORM_L2_NUMPY, 0, xt, 0.0, 0.0, False)
    np.testing.assert_allclose(y, y_ref, rtol=1e-6, atol=1e-9)
    z, _ = g.embed_device(64, 1, "left", _lib.NORM_L2_NUMPY, 0, xt, 0.0, 0.0, True)
    z_ref = oracle.whiten_embeddings(y_ref)
    assert gram_err(z, z_ref) <= 1e-5
    assert procrustes_err(z, z_ref) <= 1e-4


def test_pipelined_loop_equals_reference_stage_order(er_pair):
    """The default loop overlaps the eigensolve with the next SpMM through A(Y - 1 mu^T)T = (AY - (A1) mu^T)T
    (abi.cu embed_pipelined); with the switch off the stages run in the reference's order.  Same mathematics."""
    g, o = er_pair
    try:
        cb.set_option("pipeline_whiten", 0)
        faithful = cb.embed(g, 64, 6)
    finally:
        cb.set_option("pipeline_whiten", 1)
    piped = cb.embed(g, 64, 6)
    ref = oracle.embed(o, 64, 6)
    for got in (faithful, piped):
        assert gram_err(got, ref) <= 1e-5
        assert procrustes_err(got, ref) <= 1e-4
    assert gram_err(piped, faithful) <= 1e-5
    np.testing.assert_allclose(np.cov(piped.astype(np.float64), rowvar=False), np.eye(64), atol=1e-4)


# ------------------------------------------------------------------------------------------------ Cholesky whitening
@pytest.mark.parametrize("d", [1, 8, 32, 100, 128, 256, 512])
def test_chol_whiten_kernel(d):
    """chol_whiten.cu: T = L^-T of cov = L L^T, f64 on one SM.  T^T cov T = I, T upper triangular, equal to numpy's
    factorisation to f32 rounding; the status flag stays clear for a well-conditioned matrix and is raised for a
    rank-deficient one (the case in which the reference's 1e-10 clamp, pycleora/__init__.py:155, matters)."""
    import torch
    L = _lib.lib()
    rs = np.random.default_rng(d)
    a = rs.standard_normal((4 * d + 8, d)) * rs.uniform(0.1, 3.0, d)
    cov = np.atleast_2d(np.cov(a, rowvar=False))
    covd = torch.from_numpy(cov).cuda()
    T = torch.full((d, d), float("nan"), dtype=torch.float32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.cleora_dev_chol_whiten(covd.data_ptr(), d, T.data_ptr(), status.data_ptr(), st))
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    T64 = T.cpu().numpy().astype(np.float64)
    ref = np.linalg.inv(np.linalg.cholesky(cov)).T
    np.testing.assert_allclose(T64, ref, rtol=0, atol=1e-6 * np.max(np.abs(ref)))
    np.testing.assert_array_equal(np.tril(T64, -1), 0)
    np.testing.assert_allclose(T64.T @ cov @ T64, np.eye(d), atol=5e-6 * np.linalg.cond(cov) ** 0.5)
    if d >= 8:                                                       # singular covariance (n - 1 < d): flagged
        sing = torch.from_numpy(np.cov(rs.standard_normal((d // 2, d)), rowvar=False)).cuda()
        _lib.check(L.cleora_dev_chol_whiten(sing.data_ptr(), d, T.data_ptr(), status.data_ptr(), st))
        torch.cuda.synchronize()
        assert int(status.item()) == 1
    with pytest.raises(ValueError):
        _lib.check(L.cleora_dev_chol_whiten(covd.data_ptr(), 513, T.data_ptr(), status.data_ptr(), st))


def _with_options(fn, **opts):
    old = {k: int(_lib.lib().cleora_get_option(k.encode())) for k in opts}
    try:
        for k, v in opts.items():
            cb.set_option(k, v)
        return fn()
    finally:
        for k, v in old.items():
            cb.set_option(k, v)


@pytest.fixture(scope="module")
def er200k_pair():
    """ER 200k nodes / 2M pairs through the integer ingest; the oracle graph shares the product's CSR (itself checked
    bit-for-bit against the oracle's builder in tests/test_graph_build.py)."""
    rs = np.random.default_rng(11)
    n, e = 200_000, 2_000_000
    u, v = rs.integers(0, n, e), rs.integers(0, n, e)
    k = u != v
    g = cb.SparseMatrix.from_edge_arrays(u[k], v[k])
    rowptr, col, left, sym = g._csr()
    o = oracle.OracleGraph(rowptr, col, left, sym, g.entity_degrees, g.entity_hashes(),
                           np.zeros(g.num_entities, np.uint8), None)
    return g, o


@pytest.mark.parametrize("d", [256, 128])
def test_benchmarked_configuration_40_iterations_er20k(er_pair, d):
    """The path bench.py times -- d = 256 / 128, 40 iterations, whiten=True, default options (pipelined loop, int8
    Gram, tcgen05 apply, Cholesky-whitened inner iterations, PCA on the last) -- end to end against the oracle's
    reference-order loop (pycleora/__init__.py:109-125,963-971).  The covariance spectrum of an ER graph is nearly
    degenerate, so raw coordinates are defined only up to a rotation: the bars are the
    Procrustes-aligned iterate (<= 1e-4 of scale) and the Gram matrix (<= 1e-5).  Also: the reference stage order
    (pipeline_whiten=0) and the eigensolver-in-every-iteration variant (chol_whiten=0) against the same oracle."""
    g, o = er_pair
    ref = oracle.embed(o, d, 40)
    variants = {
        "default": lambda: cb.embed(g, d, 40),
        "reference order": lambda: _with_options(lambda: cb.embed(g, d, 40), pipeline_whiten=0),
        "eigensolver every iteration": lambda: _with_options(lambda: cb.embed(g, d, 40), chol_whiten=0),
        "reference order + eigensolver": lambda: _with_options(lambda: cb.embed(g, d, 40), pipeline_whiten=0, chol_whiten=0),
    }
    for name, fn in variants.items():
        got = fn()
        assert procrustes_err(got, ref) <= 1e-4, name
        assert gram_err(got, ref) <= 1e-5, name
        c = np.cov(got.astype(np.float64), rowvar=False)
        np.testing.assert_allclose(c, np.eye(d), atol=2e-4, err_msg=name)


@pytest.mark.parametrize("d", [256, 128])
def test_benchmarked_configuration_40_iterations_er200k(er200k_pair, d):
    """Same as above at 200k nodes / 2M pairs (nnz ~ 4.2M), where the int8 Gram runs several row slices per tile and
    the tensor-core apply many tiles per CTA."""
    g, o = er200k_pair
    ref = oracle.embed(o, d, 40)
    got = cb.embed(g, d, 40)
    assert procrustes_err(got, ref) <= 1e-4
    assert gram_err(got, ref) <= 1e-5
    got_ref_order = _with_options(lambda: cb.embed(g, d, 40), pipeline_whiten=0)
    assert procrustes_err(got_ref_order, ref) <= 1e-4
    assert gram_err(got_ref_order, ref) <= 1e-5
    # the selectable kernel variants (both settings of each, whatever the default): same bars, and -- the variants being
    # re-arrangements of identical arithmetic -- the same bits
    alt = _with_options(lambda: cb.embed(g, d, 40), k3_asw=1, gram_needed_cols=1)
    base = _with_options(lambda: cb.embed(g, d, 40), k3_asw=0, gram_needed_cols=0)
    assert procrustes_err(alt, ref) <= 1e-4 and gram_err(alt, ref) <= 1e-5
    np.testing.assert_array_equal(alt, base)


def test_cholesky_fallback_on_rank_deficient_covariance(golden_dir):
    """karate d=32: n - 1 = 33 ~ d and the covariance degenerates within a few iterations -- the Cholesky guard must
    fire and the call must fall back to the eigensolver loop (same result as with chol_whiten=0)."""
    z = np.load(os.path.join(golden_dir, "embed_karate_d32_t5_w.npz"))
    g = cb.SparseMatrix.from_iterator([str(s) for s in z["lines"]], str(z["columns"]))
    got = cb.embed(g, 32, 12)
    assert np.all(np.isfinite(got))
    lines = [f"a{i} a{(i + 1) % 9}" for i in range(9)]               # n = 9 < d = 32: singular from the first iteration
    gs = cb.SparseMatrix.from_iterator(lines, "complex::reflexive::n")
    count = _lib.lib().cleora_kernel_launch_count
    c0 = count()
    a = cb.embed(gs, 32, 4)
    c1 = count()
    b = _with_options(lambda: cb.embed(gs, 32, 4), chol_whiten=0)
    c2 = count()
    assert np.all(np.isfinite(a)) and np.all(np.isfinite(b)) and a.shape == b.shape == (9, 32)
    assert c1 - c0 > c2 - c1                                          # flagged attempt + the eigensolver loop again


# ------------------------------------------------------------------------------------------------ full-size properties
@pytest.fixture(scope="module")
def big_graph():
    """200k nodes / 4M pairs, ER, integer ingest (no oracle at this size; size-independent properties instead)."""
    rs = np.random.default_rng(1)
    n, e = 200_000, 4_000_000
    u, v = rs.integers(0, n, e), rs.integers(0, n, e)
    k = u != v
    return cb.SparseMatrix.from_edge_arrays(u[k], v[k])


def test_full_size_properties(big_graph):
    g = big_graph
    n, d = g.num_entities, 256
    ones = np.full((n, d), 0.5, np.float32)
    out = g.left_markov_propagate(ones)
    np.testing.assert_allclose(out, 0.5, rtol=2e-6)                  # rows of the left Markov operator sum to 1
    rs = np.random.default_rng(2)
    x, y = (rs.standard_normal((n, d)).astype(np.float32) for _ in range(2))
    ax, ay, axy = g.left_markov_propagate(x), g.left_markov_propagate(y), g.left_markov_propagate(x + y)
    assert scale_rel_err(axy, ax + ay) <= 1e-5                       # linearity
    emb = g.embed_fast(d, 3)
    np.testing.assert_allclose(np.linalg.norm(emb.astype(np.float64), axis=1), 1.0, rtol=1e-5)
    # symmetric operator: <x, S y> == <S x, y>
    sx, sy = g.symmetric_markov_propagate(x), g.symmetric_markov_propagate(y)
    a, b = np.vdot(x.astype(np.float64), sy.astype(np.float64)), np.vdot(sx.astype(np.float64), y.astype(np.float64))
    assert abs(a - b) <= 1e-6 * abs(a)
    w = cb.embed(g, 128, 2)
    w64 = w.astype(np.float64)
    np.testing.assert_allclose(w64.mean(0), 0, atol=1e-5)
    np.testing.assert_allclose(np.cov(w64, rowvar=False), np.eye(128), atol=2e-3)


def test_launches_are_counted_and_timings_reported(er_pair):
    g, _ = er_pair
    before = _lib.lib().cleora_kernel_launch_count()
    t = np.zeros(8)
    g.embed_device(64, 3, timings=t)
    assert _lib.lib().cleora_kernel_launch_count() - before >= 3 * 4
    assert t[2] > 0 and t[3] > 0 and t[5] > 0
