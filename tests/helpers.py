"""Shared test inputs (no reference files are read at test time)."""
import numpy as np

# Zachary's karate club as `pycleora/datasets.py:283-331` lists it (78 "u v" lines, public dataset).
_KARATE_ADJ = {
    0: [1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13, 17, 19, 21, 31], 1: [2, 3, 7, 13, 17, 19, 21, 30],
    2: [3, 7, 8, 9, 13, 27, 28, 32], 3: [7, 12, 13], 4: [6, 10], 5: [6, 10, 16], 6: [16], 8: [30, 32, 33],
    9: [33], 13: [33], 14: [32, 33], 15: [32, 33], 18: [32, 33], 19: [33], 20: [32, 33], 22: [32, 33],
    23: [25, 27, 29, 32, 33], 24: [25, 27, 31], 25: [31], 26: [29, 33], 27: [33], 28: [31, 33], 29: [32, 33],
    30: [32, 33], 31: [32, 33], 32: [33],
}
KARATE_EDGES = [f"{u} {v}" for u, vs in _KARATE_ADJ.items() for v in vs]
KARATE_COLUMNS = "complex::reflexive::member"
assert len(KARATE_EDGES) == 78


def er_lines(n, e, seed):
    rs = np.random.default_rng(seed)
    u = rs.integers(0, n, size=e)
    v = rs.integers(0, n, size=e)
    keep = u != v
    return [f"{a} {b}" for a, b in zip(u[keep], v[keep])]


def hyper_lines(n_lines, vocab, seed, kmax=6, two_col=False):
    rs = np.random.default_rng(seed)
    out = []
    for _ in range(n_lines):
        k = int(rs.integers(1, kmax + 1))
        a = " ".join(str(int(t)) for t in rs.integers(0, vocab, size=k))
        if two_col:
            k2 = int(rs.integers(1, kmax + 1))
            b = " ".join("p" + str(int(t)) for t in rs.integers(0, vocab, size=k2))
            out.append(a + "\t" + b)
        else:
            out.append(a)
    return out


def scale_rel_err(a, b):
    """max |a-b| / max |b| -- error relative to the matrix scale (used where per-element relative error is
    ill-defined because elements cancel to ~0)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def elem_rel_err(a, b, floor=0.0):
    """per-element |a-b| / max(|b|, floor)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor if floor > 0 else 1e-300)


def procrustes_err(a, b):
    """|aQ - b| / |b| with Q = argmin over orthogonal matrices."""
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    u, _, vt = np.linalg.svd(a64.T @ b64)
    return float(np.max(np.abs(a64 @ (u @ vt) - b64)) / np.max(np.abs(b64)))


def gram_err(a, b, pairs=20000, seed=0):
    """max |<a_i, a_j> - <b_i, b_j>| over sampled row pairs, relative to the largest reference inner product."""
    rs = np.random.default_rng(seed)
    i, j = rs.integers(0, a.shape[0], pairs), rs.integers(0, a.shape[0], pairs)
    ga = np.einsum("ij,ij->i", a[i].astype(np.float64), a[j].astype(np.float64))
    gb = np.einsum("ij,ij->i", b[i].astype(np.float64), b[j].astype(np.float64))
    return float(np.max(np.abs(ga - gb)) / np.max(np.abs(gb)))
