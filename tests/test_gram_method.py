"""The arithmetic of the integer Gram kernel (cleora_b200/csrc/gram_tc.cu), restated with numpy / Python integers:
quantise -> centre by an integer -> four byte planes -> seven weight-group sums -> exact combination.  This pins the
METHOD's claims (exactness, overflow bounds, the centre-correction formula) on the CPU; the kernel itself is checked
against the oracle on the GPU (tests/test_gpu_parity.py::test_integer_gram_on_tensor_cores_is_exact)."""
import math

import numpy as np
import pytest


def quant_params(x, mean):
    """quant_params_kernel: 2^e with (max|x| + max|mean|) * 2^e < 2^30."""
    bound = float(np.max(np.abs(x))) + float(np.max(np.abs(mean)))
    e = 30 - math.frexp(bound)[1] if bound > 0 else 30
    return max(-60, min(e, 60))


def byte_planes(q):
    """q (int64 holding int32 values) -> b0..b2 unsigned, b3 signed with q = sum 2^(8k) b_k."""
    u = q.astype(np.int64) & 0xFFFFFFFF
    planes = [(u >> (8 * k)) & 0xFF for k in range(4)]
    planes[3] = np.where(planes[3] >= 128, planes[3] - 256, planes[3])
    return planes


def integer_gram(x, mean):
    n, d = x.shape
    e = quant_params(x, mean)
    scale = np.float32(2.0 ** e)
    m = np.rint(mean * float(scale)).astype(np.int64)
    q = np.rint((x * scale).astype(np.float64)).astype(np.int64) - m            # __float2int_rn(v * scale) - m_j
    assert np.all(np.abs(q) < 2 ** 31)
    planes = byte_planes(q)
    np.testing.assert_array_equal(sum(p << (8 * k) for k, p in enumerate(planes)), q)      # the split is error-free
    groups = [np.zeros((d, d), dtype=object) for _ in range(7)]                  # Python ints: no overflow anywhere
    for k in range(4):
        for l in range(4):
            groups[k + l] += planes[k].astype(object).T @ planes[l].astype(object)
    Q = sum(g * (1 << (8 * s)) for s, g in enumerate(groups))
    S = q.astype(object).sum(axis=0)
    delta = np.array([mean[j] * 2.0 ** e - float(m[j]) for j in range(d)])
    cov = np.empty((d, d))
    for i in range(d):
        for j in range(d):
            acc = float(Q[i, j]) - float(S[i]) * delta[j] - delta[i] * float(S[j]) + n * delta[i] * delta[j]
            cov[i, j] = math.ldexp(acc, -2 * e)
    return cov, groups, q


@pytest.mark.parametrize("seed", [0, 1])
def test_plane_split_and_combination_reproduce_the_f64_gram(seed):
    rs = np.random.default_rng(seed)
    x = rs.standard_normal((700, 12)).astype(np.float32) * np.linspace(0.2, 2.0, 12, dtype=np.float32) + 0.3
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    mean = x.astype(np.float64).mean(0)
    cov, groups, q = integer_gram(x, mean)
    xc = x.astype(np.float64) - mean
    ref = xc.T @ xc
    assert np.max(np.abs(cov - ref)) / np.max(np.abs(ref)) < 1e-9        # only the 2^-e input grid separates them
    np.testing.assert_array_equal(cov, cov.T)
    for g in groups:                                                     # each weight group is symmetric: mirrored tiles
        assert all(g[i, j] == g[j, i] for i in range(12) for j in range(12))


def test_exact_on_grid_inputs():
    """Inputs that sit on the quantisation grid lose nothing: the result is the exact rational Gram."""
    rs = np.random.default_rng(3)
    x = (rs.integers(-2 ** 20, 2 ** 20, size=(300, 8)) / 2.0 ** 22).astype(np.float32)
    mean = x.astype(np.float64).mean(0)
    cov, _, _ = integer_gram(x, mean)
    xi = (x.astype(np.float64) * 2 ** 22).astype(np.int64).astype(object)
    n = x.shape[0]
    exact = (xi.T @ xi) * n - np.outer(xi.sum(0), xi.sum(0))             # n * sum (x-mean)(x-mean)^T in 2^-44 units
    ref = np.array([[float(v) for v in row] for row in exact]) / n / 2.0 ** 44
    assert np.max(np.abs(cov - ref)) <= 1e-15 * np.max(np.abs(ref)) + 1e-300


def test_accumulator_bounds():
    """int32 TMEM accumulators are drained every 192 stages of 32 rows; a weight group sums at most 4 plane products."""
    rows = 192 * 32
    assert 4 * 255 * 255 * rows < 2 ** 31
    # int64 global accumulators: n up to 2^31 rows
    assert 4 * 255 * 255 * 2 ** 31 < 2 ** 63
