"""The method behind the Cholesky-whitened inner iterations (cleora_b200/csrc/chol_whiten.cu), pinned on the CPU:

1. the blocked right-looking factorisation of the augmented matrix [C | I] with the kernel's own block/panel index
   maps (restated with numpy) yields T = L^-T with T^T C T = I, an upper-triangular T, and trace(C^-1) as the guard;
2. the loop-level claim: replacing the PCA whitening of iterations 1..T-1 by ANY exact whitening
   and keeping the PCA transform on the last iteration reproduces the reference's final iterate (the body of the loop
   is equivariant under orthogonal right-multiplication).  Checked with the oracle's own stages on the karate graph
   (raw per-element agreement after sign alignment) and on an ER graph (Procrustes / Gram).
The kernel itself is compared with the oracle on the GPU (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

import oracle
from tests.helpers import KARATE_COLUMNS, KARATE_EDGES, er_lines

NB = 32


def chol_whiten_blocked(cov):
    """Mirror of chol_whiten_kernel: phases P1 / P2 / P3 with the same panel column maps.  Returns (T f32, trace)."""
    d = cov.shape[0]
    dp = (d + NB - 1) // NB * NB
    nb, ldm = dp // NB, 2 * dp
    M = np.zeros((dp, ldm))
    M[:d, :d] = cov
    for r in range(d, dp):
        M[r, r] = 1.0
    M[:, dp:] = np.eye(dp)
    bad = False
    for k in range(nb):
        blk = M[k * NB:(k + 1) * NB, k * NB:(k + 1) * NB].copy()
        # P1: unblocked Cholesky (lower part of each row), then the inverse of the factor
        L = np.zeros((NB, NB))
        a = blk.copy()
        for j in range(NB):
            ajj = a[j, j]
            if not ajj > 0:
                bad, ajj = True, 1.0
            inv = 1.0 / np.sqrt(ajj)
            lj = a[:, j] * inv
            L[j:, j] = lj[j:]
            for c in range(j + 1, NB):
                a[:, c] -= lj * lj[c]
        Li = np.linalg.inv(L)                          # the kernel uses forward substitution, column per lane
        # P2: panel = inv(L_kk) * row block k, over [C blocks k+1.. | identity blocks 0..k]
        n_c = nb - k - 1
        gcols = []
        for pc in range(nb):
            base = dp + (pc - n_c) * NB if pc >= n_c else (k + 1 + pc) * NB
            gcols.append(base)
        P = np.zeros((NB, dp))
        for pc, base in enumerate(gcols):
            P[:, pc * NB:(pc + 1) * NB] = np.tril(Li) @ M[k * NB:(k + 1) * NB, base:base + NB]
            if pc >= n_c:
                M[k * NB:(k + 1) * NB, base:base + NB] = P[:, pc * NB:(pc + 1) * NB]
        # P3: row blocks i > k
        for i in range(k + 1, nb):
            ib = i - k - 1
            U = P[:, ib * NB:(ib + 1) * NB]
            for pc in range(ib, nb):
                base = gcols[pc]
                M[i * NB:(i + 1) * NB, base:base + NB] -= U.T @ P[:, pc * NB:(pc + 1) * NB]
    Z = M[:d, dp:dp + d]                               # rows of L^-1
    T = np.triu(Z.T)
    return T.astype(np.float32), float(np.sum(np.tril(Z) ** 2)), bad


@pytest.mark.parametrize("d", [1, 8, 32, 33, 100, 256])
def test_blocked_factorisation_whitens(d):
    rs = np.random.default_rng(d)
    a = rs.standard_normal((4 * d + 8, d)) * rs.uniform(0.1, 3.0, d)
    cov = np.cov(a, rowvar=False).reshape(d, d)
    T, trace, bad = chol_whiten_blocked(cov)
    assert not bad
    T64 = T.astype(np.float64)
    np.testing.assert_allclose(T64.T @ cov @ T64, np.eye(d), atol=5e-6 * np.linalg.cond(cov) ** 0.5)
    np.testing.assert_array_equal(np.tril(T, -1), 0)                       # upper triangular
    ref = np.linalg.inv(np.linalg.cholesky(cov)).T
    np.testing.assert_allclose(T64, ref, rtol=0, atol=1e-6 * np.max(np.abs(ref)))
    np.testing.assert_allclose(trace, np.trace(np.linalg.inv(cov)), rtol=1e-9)
    assert 1.0 / np.linalg.eigvalsh(cov)[0] <= trace * (1 + 1e-12)        # the guard: trace(C^-1) >= 1/lambda_min


def test_guard_flags_rank_deficient_covariance():
    rs = np.random.default_rng(0)
    a = rs.standard_normal((20, 32))                                       # n - 1 < d: singular covariance
    cov = np.cov(a, rowvar=False)
    _, trace, bad = chol_whiten_blocked(cov)
    assert bad or not trace <= 1e8


def _loop(og, d, iters, inner):
    """The reference loop (oracle stages); `inner(y)` whitens iterations 0..iters-2, the last one is PCA."""
    x = oracle.init_matrix(og.hashes, d, 0)
    for it in range(iters):
        y = oracle.normalize(oracle.spmm(og, x), "l2")
        x = oracle.whiten_embeddings(y) if it == iters - 1 else inner(y)
    return x


def _chol_whiten(y):
    mean, cov = oracle.whiten_stats(y)
    T, trace, bad = chol_whiten_blocked(cov)
    assert not bad and trace <= 1e8
    return (y - mean.astype(np.float32)) @ T


def test_inner_cholesky_whitening_reproduces_reference_iterate_karate():
    og = oracle.build_graph(KARATE_EDGES, KARATE_COLUMNS)
    ref = oracle.embed(og, 8, 40)                                          # PCA whitening in every iteration
    got = _loop(og, 8, 40, _chol_whiten)
    sign = np.sign(np.sum(got * ref, axis=0))
    assert np.max(np.abs(got * sign - ref)) <= 1e-4 * np.max(np.abs(ref))


def test_inner_cholesky_whitening_er_graph_procrustes_and_gram():
    from tests.helpers import gram_err, procrustes_err
    og = oracle.build_graph(er_lines(3000, 30000, 11), "complex::reflexive::node")
    ref = oracle.embed(og, 64, 12)
    got = _loop(og, 64, 12, _chol_whiten)
    assert procrustes_err(got, ref) <= 1e-4
    assert gram_err(got, ref) <= 1e-5
