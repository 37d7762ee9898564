"""Callers that share the loop against outputs of the unmodified reference Python
(tests/golden/callers.npz, written by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import cleora_b200 as cb
from tests.helpers import KARATE_EDGES, scale_rel_err


@pytest.fixture(scope="module")
def fx(golden_dir):
    return np.load(os.path.join(golden_dir, "callers.npz"))


def _signed(got, ref):
    s = np.sign(np.sum(got * ref, axis=0))          # eigenvector sign is an eigensolver convention
    s[s == 0] = 1
    return got * s


def _gram(a):
    a = a.astype(np.float64)
    return a @ a.T


# ------------------------------------------------------------------------------------------------ host logic (no GPU)
def test_argument_errors_match_the_reference():
    g = cb.SparseMatrix.from_iterator(KARATE_EDGES, "complex::reflexive::member")
    with pytest.raises(ValueError, match="non-empty dict"):
        cb.embed_with_node_features(g, {})
    for bad in ([], [0, 3], [2, -1], [1.5]):
        with pytest.raises(ValueError, match="positive integers"):
            cb.embed_multiscale(g, 8, scales=bad)
    with pytest.raises(ValueError, match="Unknown propagation"):
        cb.embed_multiscale(g, 8, scales=[1], propagation="up")
    with pytest.raises(ValueError, match="rows but graph has"):
        cb.embed_inductive(g, np.zeros((3, 8), np.float32), KARATE_EDGES, [], "complex::reflexive::member")


def test_update_graph_keeps_first_appearance_order():
    old, new = ["a b", "b c"], ["d a", "c e"]
    g = cb.update_graph(old, new, "complex::reflexive::x")
    assert g.entity_ids == ["a", "b", "c", "d", "e"]
    whole = cb.SparseMatrix.from_iterator(old + new, "complex::reflexive::x")
    for got, ref in zip(g.to_sparse_csr(), whole.to_sparse_csr()):
        np.testing.assert_array_equal(got, ref)


@pytest.fixture
def oracle_device(monkeypatch):
    """Test double: the two device entry points the callers use are served by the CPU oracle, so their host logic
    (feature blending, tap scheduling, warm-start row mapping, RNG use) is checked bit for bit without a GPU."""
    import oracle
    from cleora_b200 import _lib, callers
    names = {v: k for k, v in cb._DEVICE_NORMS.items()}
    graphs = {}

    def og(sm):
        return graphs[id(sm)]

    real_from_iterator = cb.SparseMatrix.from_iterator

    def from_iterator(lines, columns, hyperedge_trim_n=16, num_workers=None):
        lines = list(lines)
        sm = real_from_iterator(lines, columns, hyperedge_trim_n)
        graphs[id(sm)] = oracle.build_graph(lines, columns, hyperedge_trim_n)
        sm._keep = graphs                                    # keep ids stable for the test's lifetime
        return sm

    def embed_device(self, feature_dim, num_iterations, propagation="left", normalization=_lib.NORM_L2_NUMPY, seed=0,
                     initial_embeddings=None, residual_weight=0.0, convergence_threshold=0.0, whiten=True, out=None,
                     timings=None):
        x = initial_embeddings if initial_embeddings is not None else oracle.init_matrix(og(self).hashes, feature_dim, seed)
        res = oracle.embed(og(self), x.shape[1], num_iterations, propagation, names[normalization], seed, x, None,
                           residual_weight, convergence_threshold, whiten)
        return res, num_iterations

    monkeypatch.setattr(cb.SparseMatrix, "from_iterator", staticmethod(from_iterator))
    monkeypatch.setattr(callers.SparseMatrix, "from_iterator", staticmethod(from_iterator))
    monkeypatch.setattr(cb.SparseMatrix, "embed_device", embed_device)
    monkeypatch.setattr(cb.SparseMatrix, "initialize_deterministically",
                        lambda self, d, seed=0: oracle.init_matrix(og(self).hashes, d, seed))
    return cb


def test_callers_host_logic_bit_exact_with_oracle_double(fx, oracle_device):
    lines, cols = [str(s) for s in fx["lines"]], str(fx["columns"])
    g = cb.SparseMatrix.from_iterator(lines, cols)
    np.testing.assert_array_equal(cb.embed_multiscale(g, feature_dim=8, scales=[4, 2, 5]), fx["multiscale_w"])
    np.testing.assert_array_equal(
        cb.embed_multiscale(g, feature_dim=16, scales=[3, 9], whiten=False, propagation="symmetric"), fx["multiscale_now"])
    feats = {str(k): v for k, v in zip(fx["feat_ids"], fx["feat_vals"])}
    np.testing.assert_array_equal(cb.embed_with_node_features(g, feats, num_iterations=4, feature_weight=0.3),
                                  fx["node_features"])
    old, new = [str(s) for s in fx["old_lines"]], [str(s) for s in fx["new_lines"]]
    g_old = cb.SparseMatrix.from_iterator(old, cols)
    np.random.seed(int(fx["inductive_seed"]))
    g_new, got = cb.embed_inductive(g_old, fx["trained"], old, new, cols, num_iterations=3)
    assert g_new.entity_ids == [str(s) for s in fx["inductive_ids"]]
    np.testing.assert_array_equal(got, fx["inductive"])


# ------------------------------------------------------------------------------------------------ device parity
@pytest.mark.gpu
def test_multiscale_taps_whitened(fx):
    g = cb.SparseMatrix.from_iterator([str(s) for s in fx["lines"]], str(fx["columns"]))
    got, ref = cb.embed_multiscale(g, feature_dim=8, scales=[4, 2, 5]), fx["multiscale_w"]
    assert got.shape == ref.shape == (34, 24)
    for k in range(3):                                # taps sorted by depth: 2, 4, 5
        blk = slice(8 * k, 8 * k + 8)
        assert scale_rel_err(_signed(got[:, blk], ref[:, blk]), ref[:, blk]) <= 1e-4
        np.testing.assert_allclose(_gram(got[:, blk]), _gram(ref[:, blk]), atol=1e-3)
    # the deepest tap is what embed() returns for that depth
    np.testing.assert_allclose(_gram(got[:, 16:]), _gram(cb.embed(g, 8, 5)), atol=1e-3)


@pytest.mark.gpu
def test_multiscale_taps_unwhitened_symmetric(fx):
    g = cb.SparseMatrix.from_iterator([str(s) for s in fx["lines"]], str(fx["columns"]))
    got = cb.embed_multiscale(g, feature_dim=16, scales=[3, 9], whiten=False, propagation="symmetric")
    assert scale_rel_err(got, fx["multiscale_now"]) <= 1e-5
    assert got.shape == (34, 32)
    dup = cb.embed_multiscale(g, feature_dim=16, scales=[3, 3], whiten=False, propagation="symmetric")
    np.testing.assert_array_equal(dup[:, :16], dup[:, 16:])
    np.testing.assert_array_equal(dup[:, :16], got[:, :16])


@pytest.mark.gpu
def test_node_feature_start(fx):
    g = cb.SparseMatrix.from_iterator([str(s) for s in fx["lines"]], str(fx["columns"]))
    feats = {str(k): v for k, v in zip(fx["feat_ids"], fx["feat_vals"])}
    got, ref = cb.embed_with_node_features(g, feats, num_iterations=4, feature_weight=0.3), fx["node_features"]
    assert scale_rel_err(_signed(got, ref), ref) <= 1e-4
    feats["1"] = np.zeros(5, np.float32)
    with pytest.raises(ValueError, match="has dimension 5, expected 8"):
        cb.embed_with_node_features(g, feats)


@pytest.mark.gpu
def test_inductive_warm_start(fx):
    cols = str(fx["columns"])
    old, new = [str(s) for s in fx["old_lines"]], [str(s) for s in fx["new_lines"]]
    g_old = cb.SparseMatrix.from_iterator(old, cols)
    np.random.seed(int(fx["inductive_seed"]))
    g_new, got = cb.embed_inductive(g_old, fx["trained"], old, new, cols, num_iterations=3)
    assert g_new.entity_ids == [str(s) for s in fx["inductive_ids"]]
    ref = fx["inductive"]
    assert scale_rel_err(_signed(got, ref), ref) <= 1e-4
    np.testing.assert_allclose(_gram(got), _gram(ref), atol=1e-3)
