"""Host logic: the product's C++ graph builder and the SparseMatrix mirror against the oracle -- bit-exact CSR,
entity order, hashes, error behaviour -- no GPU needed."""
import pickle

import numpy as np
import pytest

import cleora_b200 as cb
import oracle
from tests.helpers import KARATE_COLUMNS, KARATE_EDGES, er_lines, hyper_lines


def assert_same_graph(g: cb.SparseMatrix, o: oracle.OracleGraph):
    rowptr, col, left, sym = g._csr()
    np.testing.assert_array_equal(rowptr, o.rowptr)
    np.testing.assert_array_equal(col, o.col)
    np.testing.assert_array_equal(left.view(np.uint32), o.left.view(np.uint32))       # bit-exact f32
    np.testing.assert_array_equal(sym.view(np.uint32), o.sym.view(np.uint32))
    np.testing.assert_array_equal(g.entity_degrees.view(np.uint32), o.row_sum.view(np.uint32))
    np.testing.assert_array_equal(g.entity_hashes(), o.hashes)
    assert g.entity_ids == o.entity_ids
    assert (g.num_entities, g.num_edges) == (o.n, o.nnz)


CASES = [
    ("karate", KARATE_EDGES, KARATE_COLUMNS, 16),
    ("er", er_lines(3000, 20000, 5), "complex::reflexive::node", 16),
    ("hyper_trim", hyper_lines(2000, 400, 3, kmax=24), "complex::reflexive::p", 8),
    ("hyper_notrim", hyper_lines(2000, 400, 9, kmax=12), "reflexive::complex::p", 16),
    ("two_col", hyper_lines(2000, 300, 4, kmax=6, two_col=True), "complex::u complex::p", 16),
    ("two_col_trim", hyper_lines(800, 300, 6, kmax=20, two_col=True), "complex::u complex::p", 4),
    ("simple_cols", [f"u{i % 17}\tp{i % 29}" for i in range(500)], "user product", 16),
    ("mixed", [f"u{i % 17},p{i % 29} p{i % 5}" for i in range(500)], "user complex::product", 16),
    ("ragged", ["a b", "", "   ", "a", "x\ty", "b  c", "é ü 😀", "a b c d e f g h i j k l m n o p q r s t"],
     "complex::reflexive::n", 16),
    ("unicode_ws", [" a b　", " c d "], "complex::reflexive::n", 16),
]


@pytest.mark.parametrize("name,lines,columns,trim", CASES, ids=[c[0] for c in CASES])
def test_builder_matches_oracle_bit_exact(name, lines, columns, trim):
    assert_same_graph(cb.SparseMatrix.from_iterator(iter(lines), columns, hyperedge_trim_n=trim),
                      oracle.build_graph(lines, columns, trim))


def test_snapshot_graphs(golden_dir):
    z = np.load(f"{golden_dir}/snapshots.npz")
    for tag, nnz in (("01", 1920), ("02", 5466)):
        lines = [str(s) for s in z[f"lines_{tag}"]]
        g = cb.SparseMatrix.from_iterator(lines, str(z[f"columns_{tag}"]))
        assert (g.num_entities, g.num_edges) == (100, nnz)
        assert_same_graph(g, oracle.build_graph(lines, str(z[f"columns_{tag}"])))


def test_integer_ingest_equals_string_path():
    rs = np.random.default_rng(1)
    u, v = rs.integers(0, 5000, 40000), rs.integers(0, 5000, 40000)      # includes u == v pairs and duplicates
    gp = cb.SparseMatrix.from_edge_arrays(u, v, "node")
    assert_same_graph(gp, oracle.build_graph([f"{a} {b}" for a, b in zip(u, v)], "complex::reflexive::node"))
    assert repr(gp) == "SparseMatrix(entities={}, edges={}, columns=('node', 'node'))".format(gp.num_entities, gp.num_edges)


def test_oracle_integer_ingest_equals_its_string_path():
    """oracle.graph_from_pairs (used by bench.py's reference arm, so that the CPU baseline loads nothing of the
    product) against the oracle's own line parser / hyperedge expansion, and against the product's builder."""
    rs = np.random.default_rng(2)
    u, v = rs.integers(0, 3000, 30000), rs.integers(0, 3000, 30000)      # includes u == v pairs and duplicates
    op = oracle.graph_from_pairs(u, v, "node")
    ol = oracle.build_graph([f"{a} {b}" for a, b in zip(u, v)], "complex::reflexive::node")
    for name in ("rowptr", "col", "left", "sym", "row_sum", "hashes", "column_ids"):
        np.testing.assert_array_equal(getattr(op, name), getattr(ol, name), err_msg=name)
    op.entity_ids = ol.entity_ids
    assert_same_graph(cb.SparseMatrix.from_edge_arrays(u, v, "node"), op)
    empty = oracle.graph_from_pairs(np.zeros(0, np.uint32), np.zeros(0, np.uint32))
    assert empty.n == 0 and empty.nnz == 0


def test_integer_hyperedge_ingest_equals_string_path():
    rs = np.random.default_rng(4)
    k = np.minimum(1 + rs.poisson(3.0, 4000), 24)                            # some lines above the trim limit of 16
    offsets = np.zeros(len(k) + 1, np.int64)
    np.cumsum(k, out=offsets[1:])
    members = rs.integers(0, 900, int(offsets[-1])).astype(np.uint32)
    lines = [" ".join(str(int(t)) for t in members[offsets[i]:offsets[i + 1]]) for i in range(len(k))]
    for trim in (16, 4):
        gh = cb.SparseMatrix.from_hyperedge_arrays(members, offsets, "complex::reflexive::product", trim)
        assert_same_graph(gh, oracle.build_graph(lines, "complex::reflexive::product", trim))
    with pytest.raises(ValueError):
        cb.SparseMatrix.from_hyperedge_arrays(members, offsets[:-1])


def test_from_files(tmp_path):
    p1, p2 = tmp_path / "a.tsv", tmp_path / "b.txt"
    p1.write_text("u1\tp1 p2\n\nu2\tp2\r\n")
    p2.write_text("u1\tp3\n")
    g = cb.SparseMatrix.from_files([str(p1), str(p2)], "user complex::product")
    assert_same_graph(g, oracle.build_graph(["u1\tp1 p2", "u2\tp2", "u1\tp3"], "user complex::product"))
    with pytest.raises(ValueError, match="Unsupported file format"):
        cb.SparseMatrix.from_files([str(tmp_path / "x.json")], "a b")
    with pytest.raises(ValueError, match="At least one file path"):
        cb.SparseMatrix.from_files([], "a b")


@pytest.mark.parametrize("columns,msg", [
    ("a b c", "More than one relation"), ("a", "More than one relation"),
    ("complex::reflexive::a b", "More than one relation"), ("reflexive::a", "REFLEXIVE but NOT COMPLEX"),
    ("foo::a b", "Unrecognized column field modifier: foo"),
])
def test_bad_columns_raise_value_error(columns, msg):
    with pytest.raises(ValueError, match=msg):
        cb.SparseMatrix.from_iterator(["x y"], columns)


def test_api_surface_and_errors():
    g = cb.SparseMatrix.from_iterator(KARATE_EDGES, KARATE_COLUMNS)
    assert len(g) == g.num_entities == 34 and g.num_edges == 190
    assert g.get_entity_index("33") == g.entity_ids.index("33")
    assert g.get_entity_indices(["0", "5"]) == [0, g.entity_ids.index("5")]
    with pytest.raises(ValueError, match="Entity 'nope' not found"):
        g.get_entity_index("nope")
    nb = g.get_neighbors("4")
    assert [n for n, _ in nb] == sorted([n for n, _ in nb], key=lambda s: g.entity_ids.index(s))
    assert abs(sum(v for _, v in nb) - 1.0) < 1e-6
    rows, cols, vals, n, m = g.to_sparse_csr()
    assert rows.dtype == np.uint32 and cols.dtype == np.uint32 and vals.dtype == np.float32 and (n, m) == (34, 34)
    assert rows.shape == (190,) and np.all(np.diff(rows.astype(np.int64)) >= 0)
    with pytest.raises(ValueError, match="Unknown markov_type"):
        g.to_sparse_csr("right")
    # reference quirk (src/lib.rs:180-183): for a reflexive column the name maps to col_b_id == 1 -> all False
    assert not g.get_entity_column_mask("member").any()
    with pytest.raises(ValueError, match="Column name 'x' not found"):
        g.get_entity_column_mask("x")
    g2 = cb.SparseMatrix.from_iterator(["a\tb", "c\tb"], "l r")
    np.testing.assert_array_equal(g2.get_entity_column_mask("l"), [True, False, True])
    with pytest.raises(ValueError, match="cannot be constructed directly"):
        cb.SparseMatrix(1)
    with pytest.raises(ValueError, match="Iterator elements must be strings"):
        cb.SparseMatrix.from_iterator([1, 2], "complex::reflexive::n")
    with pytest.raises(TypeError):
        g.left_markov_propagate(np.zeros((34, 4), np.float64))
    with pytest.raises(TypeError):
        g.left_markov_propagate(np.zeros(34, np.float32))
    with pytest.raises(ValueError, match="Embedding matrix has 33 rows but graph has 34 entities"):
        g.left_markov_propagate(np.zeros((33, 4), np.float32))
    with pytest.raises(ValueError, match="Unknown propagation 'up'"):
        g.embed_fast(8, 2, propagation="up")
    with pytest.raises(ValueError, match="Unknown propagation type: 'up'"):
        cb.embed(g, 8, 2, propagation="up")
    with pytest.raises(ValueError, match="num_iterations must be an int or 'auto'"):
        cb.embed(g, 8, "many")
    with pytest.raises(ValueError, match="initial_embeddings has 3 rows"):
        cb.embed(g, 8, 2, initial_embeddings=np.zeros((3, 8), np.float32))


def test_pickle_roundtrip_bincode_layout():
    g = cb.SparseMatrix.from_iterator(hyper_lines(300, 80, 2, kmax=5, two_col=True), "complex::u complex::p")
    blob = g.__getstate__()
    # bincode 1.3.3 header: u8 col_a_id, u64 len + "u", u8 col_b_id, u64 len + "p", u64 n
    assert blob[:1] == b"\x00" and blob[1:9] == (1).to_bytes(8, "little") and blob[9:10] == b"u"
    assert blob[10:11] == b"\x01" and blob[19:20] == b"p"
    assert int.from_bytes(blob[20:28], "little") == g.num_entities
    g2 = pickle.loads(pickle.dumps(g))
    assert g2.__getstate__() == blob and repr(g2) == repr(g)
    np.testing.assert_array_equal(g2.get_entity_column_mask("u"), g.get_entity_column_mask("u"))
    e = cb.SparseMatrix()
    assert (e.num_entities, e.num_edges) == (0, 0)
    with pytest.raises(RuntimeError, match="Deserialization failed"):
        cb.SparseMatrix().__setstate__(b"\x00\x01")


def test_entity_ids_setter_rehashes():
    g = cb.SparseMatrix.from_iterator(["a b"], "complex::reflexive::n")
    g.entity_ids = ["x", "yy"]
    assert g.entity_ids == ["x", "yy"]
    np.testing.assert_array_equal(g.entity_hashes(), [oracle.xxh64(b"x"), oracle.xxh64(b"yy")])
    assert g.get_entity_index("yy") == 1


# ------------------------------------------------------------------------------------------------ differential fuzz
from hypothesis import given, settings, HealthCheck  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

_TOKEN = st.sampled_from(["a", "b", "c", "d", "e", "f", "g", "h", "x1", "x2", "é", "0", "00", "A", "a "])
_CELL = st.lists(_TOKEN, min_size=0, max_size=7).map(" ".join)
_LINE = st.one_of(
    _CELL,                                                        # one column
    st.tuples(_CELL, _CELL).map("\t".join),                       # two columns, tab
    st.tuples(_CELL, _CELL).map(",".join),                        # two columns, comma
    st.tuples(_CELL, _CELL, _CELL).map("\t".join),                # wrong column count for every spec below
    st.sampled_from(["", " ", "\t", ",", "a,,b", " a\tb ", "a \t b,c"]),
)
_SPEC = st.sampled_from(["complex::reflexive::n", "REFLEXIVE::Complex::n", "complex::u complex::p", "u complex::p",
                         "complex::u p", "u p"])


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(lines=st.lists(_LINE, min_size=0, max_size=40), columns=_SPEC, trim=st.sampled_from([2, 3, 16]))
def test_builder_differential_fuzz(lines, columns, trim):
    """Two independent restatements of the reference's parser / indexer / clique expansion / trimming (C oracle,
    C++ product) must agree bit for bit on arbitrary small inputs."""
    try:
        o = oracle.build_graph(lines, columns, trim)
    except ValueError as e:
        with pytest.raises(ValueError) as info:
            cb.SparseMatrix.from_iterator(iter(lines), columns, hyperedge_trim_n=trim)
        assert str(info.value) == str(e)
        return
    assert_same_graph(cb.SparseMatrix.from_iterator(iter(lines), columns, hyperedge_trim_n=trim), o)


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(pairs=st.lists(st.tuples(st.integers(0, 12), st.integers(0, 12)), min_size=0, max_size=60))
def test_integer_ingest_differential_fuzz(pairs):
    """from_edge_arrays(src, dst) == from_iterator over the lines "src dst" (duplicates, self pairs, gaps in the ids)."""
    u = np.array([p[0] for p in pairs], np.uint32)
    v = np.array([p[1] for p in pairs], np.uint32)
    g = cb.SparseMatrix.from_edge_arrays(u, v, "node")
    o = oracle.build_graph([f"{a} {b}" for a, b in pairs], "complex::reflexive::node", 16)
    assert_same_graph(g, o)
