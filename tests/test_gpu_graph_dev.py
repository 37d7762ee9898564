"""Device-side integer ingest (cleora_b200/csrc/graph_dev.cu) against the host builder and the CPU oracle: bit-exact
CSR arrays, entity order, row sums, hashes and ids -- for whole graphs and for directly built row shards -- and the
counter-based synthetic generators.  Needs a GPU (`-m gpu`)."""
import numpy as np
import pytest

import cleora_b200 as cb
import oracle
from cleora_b200 import sharded

pytestmark = pytest.mark.gpu


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, np.uint32).view(np.int32)).cuda()


def _host(t):
    return t.cpu().numpy().view(np.uint32)


def _assert_same(gd, og, n):
    rowptr, col, left, sym = gd._csr()
    np.testing.assert_array_equal(rowptr, og.rowptr)
    np.testing.assert_array_equal(col, og.col)
    np.testing.assert_array_equal(left.view(np.uint32), og.left.view(np.uint32))
    np.testing.assert_array_equal(sym.view(np.uint32), og.sym.view(np.uint32))
    np.testing.assert_array_equal(gd.entity_degrees.view(np.uint32), og.row_sum.view(np.uint32))
    np.testing.assert_array_equal(gd.entity_hashes(), og.hashes)
    assert gd.num_entities == og.n == n and gd.num_edges == og.nnz


@pytest.mark.parametrize("n_ids,n_pairs,seed", [(50, 400, 0), (5000, 60000, 1), (200000, 1500000, 2), (7, 3, 3)])
def test_device_ingest_equals_host_builder_and_oracle(n_ids, n_pairs, seed):
    rs = np.random.default_rng(seed)
    u = rs.integers(0, n_ids, n_pairs).astype(np.uint32)
    v = rs.integers(0, n_ids, n_pairs).astype(np.uint32)                  # duplicates, self pairs, ids that never occur
    u[: n_pairs // 50] = u[-(n_pairs // 50) - 1:-1]                        # more duplicates
    gd = cb.SparseMatrix.from_edge_arrays_device(_dev(u), _dev(v))
    og = oracle.graph_from_pairs(u, v)
    _assert_same(gd, og, og.n)
    gh = cb.SparseMatrix.from_edge_arrays(u, v)
    assert gd.entity_ids == gh.entity_ids
    assert gd.get_entity_index(str(int(u[0]))) == 0                        # first appearance = index 0
    # the device-resident graph runs the hot path like any other
    x = rs.standard_normal((og.n, 32)).astype(np.float32)
    np.testing.assert_array_equal(gd.left_markov_propagate(x), oracle.spmm(og, x))
    np.testing.assert_array_equal(gd.initialize_deterministically(16, 3), oracle.init_matrix(og.hashes, 16, 3))
    np.testing.assert_array_equal(gd.symmetric_markov_propagate(x), oracle.spmm(og, x, "symmetric"))


def test_device_ingest_edge_cases():
    import torch
    e = cb.SparseMatrix.from_edge_arrays_device(torch.empty(0, dtype=torch.int32, device="cuda"),
                                                torch.empty(0, dtype=torch.int32, device="cuda"))
    assert e.num_entities == 0 and e.num_edges == 0
    g = cb.SparseMatrix.from_edge_arrays_device(_dev([5, 5, 5]), _dev([5, 5, 5]))      # only self pairs
    og = oracle.graph_from_pairs(np.uint32([5, 5, 5]), np.uint32([5, 5, 5]))
    _assert_same(g, og, 1)
    hub = np.zeros(100000, np.uint32)                                                   # a star: one very long row
    leaves = np.arange(1, 100001, dtype=np.uint32)
    g2 = cb.SparseMatrix.from_edge_arrays_device(_dev(hub), _dev(leaves), want_sym=False)
    og2 = oracle.graph_from_pairs(hub, leaves)
    rowptr, col, left, _ = og2.rowptr, og2.col, og2.left, None
    r, c, lft, _ = g2._csr()[0], g2._csr()[1], g2._csr()[2], None
    np.testing.assert_array_equal(r, rowptr)
    np.testing.assert_array_equal(c, col)
    np.testing.assert_array_equal(lft, left)
    with pytest.raises(ValueError):
        g2.symmetric_markov_propagate(np.zeros((100001, 4), np.float32))                # built without symmetric values


@pytest.mark.parametrize("world", [2, 3, 8])
def test_device_built_shards_match_host_csr(world):
    rs = np.random.default_rng(world)
    n_ids, n_pairs = 30000, 400000
    w = 1.0 / np.arange(1, n_ids + 1) ** 0.7
    w /= w.sum()
    u = rs.choice(n_ids, n_pairs, p=w).astype(np.uint32)
    v = rs.integers(0, n_ids, n_pairs).astype(np.uint32)
    gh = cb.SparseMatrix.from_edge_arrays(u, v)
    rowptr, col, left, sym = gh._csr()
    n = gh.num_entities
    ud, vd = _dev(u), _dev(v)
    shards = [sharded.Shard.from_device_pairs(ud, vd, r, world, want_sym=True) for r in range(world)]
    b = shards[0].bounds
    assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
    per = np.diff(rowptr[b])
    assert per.max() <= 1.3 * per.mean() + np.diff(rowptr).max()              # balanced by entry count
    hashes = gh.entity_hashes()
    for s in shards:
        np.testing.assert_array_equal(s.bounds, b)
        assert s.n == n and s.block == int(np.diff(b).max()) and s.n_pad == s.block * world
        lr, lc, ll, ls = s.graph._csr()
        e0, e1 = rowptr[s.r0], rowptr[s.r1]
        np.testing.assert_array_equal(lr, rowptr[s.r0:s.r1 + 1] - e0)
        np.testing.assert_array_equal(lc, sharded.padded_index(col[e0:e1], b, s.block))
        np.testing.assert_array_equal(ll, left[e0:e1])
        np.testing.assert_array_equal(ls, sym[e0:e1])
        np.testing.assert_array_equal(s.graph.entity_hashes(), hashes[s.r0:s.r1])
        np.testing.assert_array_equal(s.graph.entity_degrees, gh.entity_degrees[s.r0:s.r1])


def test_synthetic_generators():
    u, v = cb.synth_pairs("er", 100000, 2000000, seed=1)
    u2, v2 = cb.synth_pairs("er", 100000, 2000000, seed=1)
    hu, hv = _host(u), _host(v)
    assert np.array_equal(hu, _host(u2)) and np.array_equal(hv, _host(v2))   # counter-based: reproducible
    assert hu.max() < 100000 and hv.max() < 100000 and np.all(hu != hv)
    cnt = np.bincount(hu, minlength=100000)
    assert abs(cnt.mean() - 20.0) < 0.01 and cnt.max() < 60                   # uniform endpoints
    u3, _ = cb.synth_pairs("er", 100000, 2000000, seed=2)
    assert not np.array_equal(hu, _host(u3))
    cu, cv = cb.synth_pairs("chunglu", 200000, 4000000, seed=5, alpha=0.833)
    hcu, hcv = _host(cu), _host(cv)
    assert hcu.max() < 200000 and np.all(hcu != hcv)
    deg = np.bincount(np.concatenate([hcu, hcv]), minlength=200000)
    assert deg.max() > 200 * np.median(deg[deg > 0])                          # heavy tail
    assert (deg > 0).mean() > 0.5
