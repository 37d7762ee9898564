#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/.  Run ONCE in the build container, where
/root/reference exists (it does not exist on the GPU box; tests only read the committed .npz files).

  python tests/golden/make_golden.py

Fixtures
--------
snapshots.npz     The reference's four insta snapshots (tests/snapshots/snapshot__tests__markov_*.snap) as int32
                  [100,32] arrays, together with the INPUTS that tests/snapshot.rs:52-117 builds them from: the
                  hyperedge lines and the [100,32] f32 embedding matrix, regenerated here by restating the test's
                  RNG (rand 0.8.5 StdRng = ChaCha12, seeded through rand_core 0.6.4's PCG32 `seed_from_u64`;
                  `Uniform::new(0.,10.)` f32 sampling from rand 0.8.5).  The oracle must map inputs -> snapshot ints
                  exactly; that pins graph build, both Markov normalisations and the SpMM order.
xxh64_kat.npz     XXH64(seed 0) known answers from the python `xxhash` package (public implementation), because
                  the reference has no test pinning twox-hash.
embed_*.npz       Outputs of the UNMODIFIED reference Python (`/root/reference/pycleora/__init__.py`: embed,
                  whiten_embeddings, _normalize, _compute_rmse) run in this container on top of a stub
                  `pycleora.pycleora.SparseMatrix` whose native methods are served by the C oracle.  They pin
                  the numpy half of the oracle (oracle.normalize / whiten_* / embed).
"""
import importlib.util
import os
import re
import struct
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


# ---------------------------------------------------------------------------------------------- StdRng restatement
def _rotl32(x, r):
    return ((x << r) | (x >> (32 - r))) & M32


class StdRng:
    """rand 0.8.5 `StdRng` (= rand_chacha 0.3.1 ChaCha12Rng) created by `SeedableRng::seed_from_u64`."""

    def __init__(self, seed_u64: int):
        state = seed_u64 & M64
        key = []
        for _ in range(8):  # rand_core 0.6.4 seed_from_u64: PCG32 output per 4-byte chunk
            state = (state * 6364136223846793005 + 11634580027462260723) & M64
            xorshifted = ((((state >> 18) ^ state) >> 27)) & M32
            rot = state >> 59
            key.append(((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & M32)
        self.key = key
        self.counter = 0
        self.buf = []

    def _block(self):
        c = self.counter
        inp = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + self.key + [c & M32, (c >> 32) & M32, 0, 0]
        x = list(inp)

        def qr(a, b, c_, d):
            x[a] = (x[a] + x[b]) & M32; x[d] = _rotl32(x[d] ^ x[a], 16)
            x[c_] = (x[c_] + x[d]) & M32; x[b] = _rotl32(x[b] ^ x[c_], 12)
            x[a] = (x[a] + x[b]) & M32; x[d] = _rotl32(x[d] ^ x[a], 8)
            x[c_] = (x[c_] + x[d]) & M32; x[b] = _rotl32(x[b] ^ x[c_], 7)

        for _ in range(6):  # 12 rounds
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        self.counter += 1
        return [(x[i] + inp[i]) & M32 for i in range(16)]

    def next_u32(self) -> int:
        if not self.buf:
            self.buf = self._block()
        return self.buf.pop(0)

    def uniform_f32(self, low: float, high: float) -> np.float32:
        """rand 0.8.5 UniformFloat<f32>::sample: value1_2 from the top 23 bits, (v - 1) * scale + low."""
        bits = (self.next_u32() >> 9) | 0x3F800000
        v12 = np.frombuffer(struct.pack("<I", bits), dtype=np.float32)[0]
        scale = np.float32(high) - np.float32(low)
        return np.float32(np.float32(v12 - np.float32(1.0)) * scale) + np.float32(low)


def snapshot_inputs(kind: str):
    """tests/snapshot.rs:52-87 (complex_complex) / :89-117 (complex_reflexive)."""
    rng = StdRng(2137)
    lines = []
    for _ in range(1000):
        if kind == "reflexive":
            a, b = rng.next_u32() % 100, rng.next_u32() % 100
            lines.append(f"{a} {b}")
        else:
            a, b, c, d = (rng.next_u32() % 100 for _ in range(4))
            lines.append(f"{a} {b}\t{c} {d}")
    emb = np.array([rng.uniform_f32(0.0, 10.0) for _ in range(100 * 32)], dtype=np.float32).reshape(100, 32)
    columns = "reflexive::complex::entity_id" if kind == "reflexive" else "complex::entity_a complex::entity_b"
    return lines, columns, emb


def parse_snap(path: str) -> np.ndarray:
    txt = open(path).read()
    body = txt.split("---", 2)[2]
    body = body[:body.index("]]") + 2]          # drop ndarray's Debug trailer (", shape=[100, 32], ...")
    vals = np.array([int(v) for v in re.findall(r"-?\d+", body)], dtype=np.int32)
    assert vals.size == 3200, (path, vals.size)
    return vals.reshape(100, 32)


def make_snapshots():
    out = {}
    for tag, kind in (("01", "reflexive"), ("02", "complex")):
        lines, columns, emb = snapshot_inputs(kind)
        out[f"lines_{tag}"] = np.array(lines)
        out[f"columns_{tag}"] = np.array(columns)
        out[f"emb_{tag}"] = emb
        for mk in ("left", "sym"):
            out[f"{mk}_{tag}"] = parse_snap(f"{REF}/tests/snapshots/snapshot__tests__markov_{mk}_{tag}.snap")
    np.savez_compressed(os.path.join(HERE, "snapshots.npz"), **out)
    print("snapshots.npz written")


def make_xxh_kat():
    import xxhash
    rs = np.random.RandomState(7)
    msgs = [b"", b"a", b"abc", b"0", b"33", b"cleora", b"message digest", b"abcdefghijklmnopqrstuvwxyz",
            "zażółć gęślą jaźń".encode(), bytes(range(256))]
    for ln in (1, 3, 4, 7, 8, 15, 16, 31, 32, 33, 63, 64, 65, 100, 1000):
        msgs.append(rs.bytes(ln))
    digests = np.array([xxhash.xxh64(m, seed=0).intdigest() for m in msgs], dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, "xxh64_kat.npz"),
                        msgs=np.array([m.hex() for m in msgs]), digests=digests)
    print("xxh64_kat.npz written")


# ---------------------------------------------------------------------------------------------- reference import
def import_reference():
    """Import /root/reference/pycleora/__init__.py unmodified, with the (unbuildable: no Rust toolchain)
    native module `pycleora.pycleora` replaced by a stub whose methods are served by the C oracle."""
    import oracle

    class SparseMatrix:  # minimal stand-in for src/lib.rs #[pymethods]
        def __init__(self):
            self._g = None

        @staticmethod
        def from_iterator(hyperedges, columns, hyperedge_trim_n=16, num_workers=None):
            sm = SparseMatrix()
            sm._g = oracle.build_graph(list(hyperedges), columns, hyperedge_trim_n)
            return sm

        entity_ids = property(lambda self: list(self._g.entity_ids))
        num_entities = property(lambda self: self._g.n)
        num_edges = property(lambda self: self._g.nnz)

        def initialize_deterministically(self, feature_dim, seed=0):
            return oracle.init_matrix(self._g.hashes, feature_dim, seed)

        def left_markov_propagate(self, x, num_workers=None):
            return oracle.spmm(self._g, x, "left")

        def symmetric_markov_propagate(self, x, num_workers=None):
            return oracle.spmm(self._g, x, "symmetric")

        def embed_fast(self, feature_dim, num_iterations, propagation="left", seed=0, residual_weight=0.0,
                       num_workers=None):
            return oracle.embed_fast(self._g, feature_dim, num_iterations, propagation, seed, residual_weight)

        def embed_fast_convergence(self, feature_dim, max_iterations, propagation="left", seed=0,
                                   residual_weight=0.0, convergence_threshold=0.0, num_workers=None):
            return oracle.embed_fast_convergence(self._g, feature_dim, max_iterations, propagation, seed,
                                                 residual_weight, convergence_threshold)

    stub = types.ModuleType("pycleora.pycleora")
    stub.SparseMatrix = SparseMatrix
    spec = importlib.util.spec_from_file_location("pycleora", f"{REF}/pycleora/__init__.py",
                                                  submodule_search_locations=[f"{REF}/pycleora"])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["pycleora"] = mod
    sys.modules["pycleora.pycleora"] = stub
    spec.loader.exec_module(mod)
    return mod, SparseMatrix


def er_lines(n, e, seed):
    rs = np.random.default_rng(seed)
    u = rs.integers(0, n, size=e)
    v = rs.integers(0, n, size=e)
    keep = u != v
    return [f"{a} {b}" for a, b in zip(u[keep], v[keep])]


def make_reference_runs(only=None):
    ref, SM = import_reference()
    sys.path.insert(0, f"{REF}")
    karate = None
    spec = importlib.util.spec_from_file_location("ref_datasets", f"{REF}/pycleora/datasets.py")
    dsm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dsm)
    karate = dsm.load_karate_club()

    cases = {
        # name: (lines, columns, kwargs for embed)
        "karate_d8_t5_w": (karate["edges"], karate["columns"], dict(feature_dim=8, num_iterations=5)),
        "karate_d32_t5_w": (karate["edges"], karate["columns"], dict(feature_dim=32, num_iterations=5)),
        "karate_d8_t40_w": (karate["edges"], karate["columns"], dict(feature_dim=8, num_iterations=40)),
        "karate_d32_t40_now": (karate["edges"], karate["columns"],
                               dict(feature_dim=32, num_iterations=40, whiten=False)),
        "karate_d16_t10_sym_res": (karate["edges"], karate["columns"],
                                   dict(feature_dim=16, num_iterations=10, propagation="symmetric",
                                        residual_weight=0.25)),
        "karate_d8_t30_conv": (karate["edges"], karate["columns"],
                               dict(feature_dim=8, num_iterations=30, convergence_threshold=0.05)),
        "karate_d16_t6_l1": (karate["edges"], karate["columns"],
                             dict(feature_dim=16, num_iterations=6, normalization="l1")),
        "er2k_d32_t5_w": (er_lines(2000, 12000, 11), "complex::reflexive::node",
                          dict(feature_dim=32, num_iterations=5)),
        "er2k_d64_t8_now": (er_lines(2000, 12000, 12), "complex::reflexive::node",
                            dict(feature_dim=64, num_iterations=8, whiten=False)),
        # normalization="spectral" (pycleora/__init__.py:951-956: row l2 norm, then U*S of the SVD = a rotation)
        "karate_d8_t6_spectral_now": (karate["edges"], karate["columns"],
                                      dict(feature_dim=8, num_iterations=6, normalization="spectral", whiten=False)),
        "karate_d8_t6_spectral_w": (karate["edges"], karate["columns"],
                                    dict(feature_dim=8, num_iterations=6, normalization="spectral")),
    }
    for name, (lines, columns, kw) in cases.items():
        if only is not None and name not in only:
            continue
        g = SM.from_iterator(iter(lines), columns)
        trace = []
        kw2 = dict(kw)
        if name.endswith("_w") and kw["num_iterations"] <= 8:
            kw2["callback"] = lambda i, e: trace.append(e.copy())   # per-iteration taps (teacher forcing)
        out = ref.embed(g, **kw2)
        save = dict(lines=np.array(lines), columns=np.array(columns), out=out,
                    kwargs=np.array(repr(kw)))
        if trace:
            save["trace"] = np.stack(trace)
        np.savez_compressed(os.path.join(HERE, f"embed_{name}.npz"), **save)
        print(f"embed_{name}.npz written: out {out.shape}")

    if only is not None:
        return
    # callers that share the loop: multiscale taps, node-feature start, inductive warm start
    g = SM.from_iterator(iter(karate["edges"]), karate["columns"])
    ids = g.entity_ids
    rs = np.random.default_rng(21)
    feats = {eid: rs.standard_normal(8).astype(np.float32) for eid in ids[::3]}
    feats["not-in-graph"] = np.zeros(8, np.float32)
    old_lines, new_lines = list(karate["edges"][:60]), list(karate["edges"][60:]) + ["34 35", "35 1"]
    g_old = SM.from_iterator(iter(old_lines), karate["columns"])
    trained = ref.embed(g_old, feature_dim=8, num_iterations=4)
    np.random.seed(1234)
    drawn = np.random.randn(36, 8)
    np.random.seed(1234)
    g_new, ind = ref.embed_inductive(g_old, trained, old_lines, new_lines, karate["columns"], num_iterations=3)
    np.savez_compressed(
        os.path.join(HERE, "callers.npz"), lines=np.array(karate["edges"]), columns=np.array(karate["columns"]),
        multiscale_w=ref.embed_multiscale(g, feature_dim=8, scales=[4, 2, 5]),
        multiscale_now=ref.embed_multiscale(g, feature_dim=16, scales=[3, 9], whiten=False, propagation="symmetric"),
        feat_ids=np.array(list(feats.keys())), feat_vals=np.stack(list(feats.values())),
        node_features=ref.embed_with_node_features(g, feats, num_iterations=4, feature_weight=0.3),
        old_lines=np.array(old_lines), new_lines=np.array(new_lines), trained=trained,
        inductive_ids=np.array(g_new.entity_ids), inductive=ind, inductive_seed=np.int64(1234),
        inductive_draw_shape=np.array(drawn.shape))
    print("callers.npz written")

    # stage-wise whitening fixture: reference whiten_embeddings + _normalize on a fixed random matrix
    rs = np.random.default_rng(5)
    x = rs.standard_normal((3000, 48)).astype(np.float32) * np.linspace(0.5, 3.0, 48, dtype=np.float32) + 0.3
    xn = ref._normalize(x, "l2")
    np.savez_compressed(os.path.join(HERE, "whiten_stage.npz"), x=x, normalized=xn,
                        whitened=ref.whiten_embeddings(xn), l1=ref._normalize(x, "l1"),
                        rmse=np.float64(ref._compute_rmse(xn, x)))
    print("whiten_stage.npz written")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    if len(sys.argv) > 1:                       # python make_golden.py <embed case> ...: (re)generate only those
        make_reference_runs(only=set(sys.argv[1:]))
        sys.exit(0)
    make_snapshots()
    make_xxh_kat()
    make_reference_runs()
