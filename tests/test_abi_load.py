"""The C-ABI library loads and exports every symbol include/cleora_b200.h declares; without a GPU the compute
entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from cleora_b200 import _lib
import cleora_b200 as cb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cleora_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(cleora_[a-z0-9_]+)\s*\(", src))
    names.discard("cleora_eigh_fn")
    return names


def test_every_declared_symbol_is_exported_and_prototyped():
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.PROTOTYPES), (declared ^ set(_lib.PROTOTYPES))


def test_version_and_hash():
    L = _lib.lib()
    assert b"sm_100a" in L.cleora_version()
    assert L.cleora_hash_entity(b"cleora", 6) == 0xABC2642DCA9956DD
    assert L.cleora_hash_entity(b"", 0) == 0xEF46DB3751D8E999


def test_no_cpu_fallback():
    L = _lib.lib()
    if L.cleora_device_count() > 0:
        pytest.skip("a GPU is present")
    g = cb.SparseMatrix.from_iterator(["a b", "b c"], "complex::reflexive::n")
    x = np.zeros((3, 8), np.float32)
    for call in (lambda: g.left_markov_propagate(x), lambda: g.initialize_deterministically(8),
                 lambda: g.embed_fast(8, 2), lambda: cb.embed(g, 8, 2), lambda: cb.whiten_embeddings(x),
                 lambda: g.l2_normalize(x)):
        with pytest.raises(RuntimeError, match="no CUDA device"):
            call()
    cb.set_devices([0, 0])                                   # the in-process multi-GPU entry point fails the same way
    try:
        for call in (lambda: cb.embed(g, 64, 2), lambda: g.embed_fast(64, 2)):
            with pytest.raises(RuntimeError, match="no CUDA device"):
                call()
    finally:
        cb.set_devices(None)


def test_product_never_imports_the_oracle():
    """The product package must not reference oracle/ anywhere."""
    pkg = os.path.join(ROOT, "cleora_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "orc_" not in txt, f


def test_multi_gpu_shape_table_and_device_list_api():
    """cleora_embed_multi_supported (which (d, n_devices) pairs the in-process multi-GPU loop takes) and the Python
    device-list plumbing around it -- host logic only, no compute."""
    from cleora_b200 import pycleora as pc
    L = _lib.lib()
    widths = (8, 16, 32, 64, 96, 128, 192, 256, 384, 512, 1024)
    for d in (8, 24, 48, 64, 96, 100, 128, 192, 256, 384, 512, 1024, 2048):
        for w in range(1, 10):
            want = 1 <= w <= 8 and d % w == 0 and d in widths and (d // w) in widths
            assert bool(L.cleora_embed_multi_supported(d, w)) == want, (d, w)
    assert L.cleora_embed_multi_supported(0, 2) == 0 and L.cleora_embed_multi_supported(-64, 2) == 0
    old = cb.get_devices()
    try:
        cb.set_devices([0, 1, 2, 3])
        assert cb.get_devices() == [0, 1, 2, 3]
        assert pc._multi_devices(256) is not None and list(pc._multi_devices(256)) == [0, 1, 2, 3]
        assert pc._multi_devices(48) is None                      # 12-float slices: no kernel -> the one-GPU path
        cb.set_devices([0])
        assert pc._multi_devices(256) is None                     # a single device is the ordinary path
        cb.set_devices(None)
        assert cb.get_devices() == [] and pc._multi_devices(256) is None
    finally:
        cb.set_devices(old)
    # argument checks that need no GPU: an empty device list is a value error before any CUDA call
    g = cb.SparseMatrix.from_iterator(["a b", "b c"], "complex::reflexive::n")
    out = np.empty((3, 64), np.float32)
    rc = L.cleora_embed_multi(g._handle(), None, 0, None, 64, 2, 0, 0, 0.0, 0.0, _lib.NORM_L2_NUMPY, 1,
                              _lib.ptr(out, _lib.c_f32p), None)
    assert rc == _lib.ERR_VALUE and b"at least one device" in L.cleora_last_error()
