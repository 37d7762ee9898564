"""The import path ``pycleora.pycleora`` (src/lib.rs:490-495): the reference's UNMODIFIED Python package
runs on top of ``shim/pycleora.py``.  The reference tree only exists in the build container, so this test assembles a
throw-away package there (reference ``pycleora/*.py`` copied into a temp directory at test time -- never into the repo --
plus the shim in place of the native module) and is skipped elsewhere.  No GPU here: the two compute entry points the
reference's ``embed()`` calls are served by the CPU oracle, which makes this a test of the plumbing -- C1 of
BASELINE.json: karate club, d=32, 5 iterations, through the reference's own ``embed()`` and CLI loader."""
import importlib
import os
import shutil
import sys

import numpy as np
import pytest

REF = "/root/reference/pycleora"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")


@pytest.fixture
def ref_pkg(tmp_path, monkeypatch):
    pkg = tmp_path / "pycleora"
    pkg.mkdir()
    for name in os.listdir(REF):
        if name.endswith((".py", ".pyi", ".typed")):
            shutil.copy(os.path.join(REF, name), pkg / name)
    shutil.copy(os.path.join(ROOT, "shim", "pycleora.py"), pkg / "pycleora.py")      # in place of pycleora.*.so
    monkeypatch.syspath_prepend(str(tmp_path))
    for m in [k for k in sys.modules if k == "pycleora" or k.startswith("pycleora.")]:
        monkeypatch.delitem(sys.modules, m)
    mod = importlib.import_module("pycleora")
    yield mod
    for m in [k for k in sys.modules if k == "pycleora" or k.startswith("pycleora.")]:
        sys.modules.pop(m, None)


def test_reference_package_imports_on_the_shim_and_runs_c1(ref_pkg, monkeypatch):
    import cleora_b200
    import oracle
    pycleora = ref_pkg
    assert pycleora.SparseMatrix is cleora_b200.SparseMatrix                   # pycleora/__init__.py:4 resolved to the shim
    assert importlib.import_module("pycleora.pycleora").SparseMatrix is cleora_b200.SparseMatrix
    # C1: the reference's own dataset loader and graph construction (cli.py:137-145 does exactly this)
    from pycleora.datasets import load_dataset
    ds = load_dataset("karate_club")
    graph = pycleora.SparseMatrix.from_iterator(iter(ds["edges"]), ds["columns"])
    assert graph.num_entities == 34 and graph.num_edges == 190
    og = oracle.build_graph(list(ds["edges"]), ds["columns"])
    assert graph.entity_ids == og.entity_ids
    rows, cols, vals, n, _ = graph.to_sparse_csr()
    np.testing.assert_array_equal(vals, og.left)
    # no GPU in this container: serve the two compute calls of the reference loop from the oracle
    SM = cleora_b200.SparseMatrix
    monkeypatch.setattr(SM, "initialize_deterministically", lambda self, d, seed=0: oracle.init_matrix(og.hashes, d, seed))
    monkeypatch.setattr(SM, "left_markov_propagate", lambda self, x, num_workers=None: oracle.spmm(og, x, "left"))
    got = pycleora.embed(graph, 32, 5)                                         # the reference's unmodified embed()
    np.testing.assert_array_equal(got, oracle.embed(og, 32, 5))                # == the oracle's restatement of it
    got2 = pycleora.embed_using_baseline_cleora(graph, 32, 5)
    np.testing.assert_array_equal(got2, got)
    # compute without a device fails loudly through the reference's call path, never silently on the CPU
    monkeypatch.undo()
    g2 = cleora_b200.SparseMatrix.from_iterator(iter(ds["edges"]), ds["columns"])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        g2.left_markov_propagate(np.zeros((34, 4), np.float32))
