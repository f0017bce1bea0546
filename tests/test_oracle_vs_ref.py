"""Pins the C restatement (oracle/gcn_oracle.c) bit-for-bit against the UNMODIFIED reference compiled
into oracle/_ref/libgcnref.so.  CPU only.  Skipped where the reference library was never built."""
import numpy as np
import pytest

from tests.util import make_dataset, write_text_dataset


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bits(a, b, what=""):
    a, b = bits(a), bits(b)
    assert a.shape == b.shape, what
    bad = np.flatnonzero(a != b)
    assert bad.size == 0, f"{what}: {bad.size} of {a.size} floats differ, first at {bad[:5]}"


def test_rng_stream_and_seed(oracle, ref):
    for seed in (0, 1, 42, 1571234567):
        oracle.init_rand_state(seed)
        ref.init_rand_state(seed)
        assert oracle.get_rand_state() == ref.get_rand_state()
        assert (oracle.rand(1000) == ref.rand(1000)).all()


def test_glorot(oracle, ref):
    for (i, o) in ((1433, 16), (16, 7), (602, 16), (3, 5)):
        oracle.init_rand_state(7)
        ref.init_rand_state(7)
        assert_bits(oracle.glorot(i, o), ref.glorot(i, o), f"glorot {i}x{o}")
        assert oracle.get_rand_state() == ref.get_rand_state()


@pytest.mark.parametrize("m,n,p", [(50, 16, 7), (33, 16, 41), (7, 5, 3), (1, 1, 1), (40, 64, 47)])
def test_matmul(oracle, ref, m, n, p):
    rng = np.random.default_rng(m * 100 + n)
    a, b, g = (rng.standard_normal(s).astype(np.float32) for s in (m * n, n * p, m * p))
    assert_bits(oracle.matmul_fw(a, b, m, n, p), ref.matmul_fw(a, b, m, n, p), "matmul fw")
    oa, ob = oracle.matmul_bw(a, b, g, m, n, p)
    ra, rb = ref.matmul_bw(a, b, g, m, n, p)
    assert_bits(oa, ra, "matmul bw A")
    assert_bits(ob, rb, "matmul bw B")


@pytest.mark.parametrize("dense", [False, True])
def test_sparse_matmul(oracle, ref, dense):
    d = make_dataset(n=120, f=40, nnz_per_row=6, dense=dense, seed=3, empty_rows=0 if dense else 4)
    p = 16
    rng = np.random.default_rng(5)
    w = rng.standard_normal(d.input_dim * p).astype(np.float32)
    g = rng.standard_normal(d.num_nodes * p).astype(np.float32)
    args = (d.feature_indptr, d.feature_indices, d.feature_value)
    assert_bits(oracle.spmm_fw(*args, w, d.num_nodes, d.input_dim, p), ref.spmm_fw(*args, w, d.num_nodes, d.input_dim, p))
    assert_bits(oracle.spmm_bw(*args, g, d.num_nodes, d.input_dim, p), ref.spmm_bw(*args, g, d.num_nodes, d.input_dim, p))


@pytest.mark.parametrize("dim", [1, 3, 7, 16, 41])
@pytest.mark.parametrize("kind", ["plain", "isolated_hub", "directed"])
def test_graphsum(oracle, ref, dim, kind):
    kw = dict(plain={}, isolated_hub=dict(isolated=9, hub=(2, 180)), directed=dict(symmetric=False))[kind]
    d = make_dataset(n=250, n_undirected=900, seed=dim, **kw)
    x = np.random.default_rng(dim).standard_normal(d.num_nodes * dim).astype(np.float32)
    for backward in (False, True):
        assert_bits(oracle.graphsum(d.graph_indptr, d.graph_indices, x, dim),
                    ref.graphsum(d.graph_indptr, d.graph_indices, x, dim, backward=backward), f"graphsum bw={backward}")


@pytest.mark.parametrize("training", [True, False])
def test_cross_entropy(oracle, ref, training):
    rng = np.random.default_rng(11)
    n, c = 300, 7
    logits = (rng.standard_normal(n * c) * 3).astype(np.float32)
    truth = rng.integers(-1, c, n).astype(np.int32)
    lo, lgo, go = oracle.cross_entropy(logits, truth, c, training)
    lr, lgr, gr = ref.cross_entropy(logits, truth, c, training)
    assert np.float32(lo).view(np.uint32) == np.float32(lr).view(np.uint32)
    assert_bits(lgo, lgr, "in-place shifted logits")
    if training:
        assert_bits(go, gr, "ce grad")


def test_cross_entropy_no_labels_is_nan(oracle, ref):
    logits = np.ones(12, np.float32)
    truth = -np.ones(4, np.int32)
    assert np.isnan(oracle.cross_entropy(logits, truth, 3, False)[0])
    assert np.isnan(ref.cross_entropy(logits, truth, 3, False)[0])


def test_relu_dropout(oracle, ref):
    rng = np.random.default_rng(2)
    x = rng.standard_normal(1000).astype(np.float32)
    x[::17] = 0.0
    x[5] = -0.0
    g = rng.standard_normal(1000).astype(np.float32)
    for training in (True, False):
        xo, mo, go = oracle.relu(x, g if training else None, training)
        xr, mr, gr = ref.relu(x, g if training else None, training)
        assert_bits(xo, xr)
        if training:
            assert (mo == mr).all()
            assert_bits(go, gr)
    for p in (0.0, 0.1, 0.5, 0.9):
        for with_grad in (True, False):
            oracle.set_rand_state(123456789, 987654321)
            ref.set_rand_state(123456789, 987654321)
            xo, mo, go = oracle.dropout(x, p, g, True, with_grad)
            xr, mr, gr = ref.dropout(x, p, g, True, with_grad)
            assert_bits(xo, xr, f"dropout p={p}")
            assert oracle.get_rand_state() == ref.get_rand_state()
            if with_grad:
                assert (mo == mr).all()
                assert_bits(go, gr)
    # eval: no draws, no change
    oracle.set_rand_state(5, 6)
    xo, _, _ = oracle.dropout(x, 0.5, None, False)
    assert_bits(xo, x)
    assert oracle.get_rand_state() == (5, 6)


def test_adam(oracle, ref):
    rng = np.random.default_rng(9)
    datas = [rng.standard_normal(500).astype(np.float32), rng.standard_normal(77).astype(np.float32)]
    steps = [[(rng.standard_normal(500) * 10.0 ** rng.integers(-6, 1)).astype(np.float32),
              rng.standard_normal(77).astype(np.float32)] for _ in range(25)]
    o = oracle.adam(datas, steps, [1, 0], 0.01, 5e-4)
    r = ref.adam(datas, steps, [1, 0], 0.01, 5e-4)
    for a, b in zip(o, r):
        assert_bits(a, b, "adam")


def test_parser(tmp_path, oracle, ref):
    d = make_dataset(n=60, f=30, c=4, n_undirected=150, nnz_per_row=5, seed=8, isolated=3, empty_rows=2)
    write_text_dataset(tmp_path, "toy", d)
    po, pr = oracle.parse(tmp_path, "toy"), ref.parse(tmp_path, "toy")
    assert po is not None and pr is not None
    for k in pr:
        if isinstance(pr[k], np.ndarray):
            assert po[k].dtype == pr[k].dtype and po[k].shape == pr[k].shape, k
            assert (po[k].view(np.uint32) == pr[k].view(np.uint32)).all(), k
        else:
            assert po[k] == pr[k], k
    # the writer round-trips the generated arrays exactly (ints) / to print precision (floats)
    assert (po["graph_indptr"] == d.graph_indptr).all() and (po["graph_indices"] == d.graph_indices).all()
    assert (po["feature_indices"] == d.feature_indices).all() and (po["label"] == d.label).all()
    assert oracle.parse(tmp_path, "missing") is None and ref.parse(tmp_path, "missing") is None


def test_parser_quirks(tmp_path, oracle, ref):
    """Accepted-input behaviour from SURVEY Appendix B: duplicates kept, parsing of a graph line stops at the
    first non-integer token, blank svmlight line -> label -1, unterminated last line dropped."""
    root = tmp_path / "data"
    root.mkdir()
    (root / "q.graph").write_text("1 2\n0 0 x 5\n\n 2   1 \n0 1")        # last line has no newline
    (root / "q.split").write_text("1\n2\n3\n0\n")
    (root / "q.svmlight").write_text("0 0:1.5 3:2\n2 1:0.25\n\n1 2:1e-3 2:7\n")
    po, pr = oracle.parse(tmp_path, "q"), ref.parse(tmp_path, "q")
    assert pr["num_nodes"] == 4 and list(pr["graph_indices"]) == [0, 1, 2, 1, 0, 0, 2, 3, 2, 1]
    assert list(pr["label"]) == [0, 2, -1, 1]
    for k in pr:
        if isinstance(pr[k], np.ndarray):
            assert (po[k].view(np.uint32) == pr[k].view(np.uint32)).all(), k
        else:
            assert po[k] == pr[k], k


@pytest.mark.parametrize("dropout", [0.0, 0.5])
def test_full_training_run(oracle, ref, dropout):
    d = make_dataset(n=300, f=50, c=6, n_undirected=1000, nnz_per_row=7, seed=21, isolated=4)
    go = oracle.gcn(d, hidden_dim=16, dropout=dropout, epochs=8, seed=77)
    gr = ref.gcn(d, hidden_dim=16, dropout=dropout, epochs=8, seed=77)
    for idx in (2, 5):
        assert_bits(go.var(idx), gr.var(idx), f"initial weights V{idx}")
    for epoch in range(8):
        to, tr = go.train_epoch(), gr.train_epoch()
        assert_bits(to, tr, f"train epoch {epoch}")
        for idx in (1, 2, 3, 4, 5, 6):
            assert_bits(go.var(idx), gr.var(idx), f"V{idx} after epoch {epoch}")
            assert_bits(go.var(idx, True), gr.var(idx, True), f"V{idx}.grad after epoch {epoch}")
        vo, vr = go.eval(2), gr.eval(2)
        assert_bits(vo, vr, f"val epoch {epoch}")
    assert_bits(go.eval(3), gr.eval(3), "test")
    go.close()
    gr.close()
