"""The C restatement (oracle/gcn_oracle.c) against the committed golden vectors that the UNMODIFIED
reference produced (tools/make_golden.py).  CPU only; needs neither /root/reference nor oracle/_ref."""
from pathlib import Path

import numpy as np
import pytest

from oracle.checker import GraphData

GOLD = Path(__file__).resolve().parent / "golden"


def same_bits(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return a.shape == b.shape and bool((a == b).all())


def load_data(z, prefix="d_"):
    dims = z[prefix + "dims"]
    return GraphData(z[prefix + "graph_indptr"], z[prefix + "graph_indices"], z[prefix + "feature_indptr"],
                     z[prefix + "feature_indices"], z[prefix + "feature_value"], z[prefix + "label"],
                     z[prefix + "split"], input_dim=int(dims[1]), output_dim=int(dims[2]))


@pytest.fixture(scope="module")
def ops():
    return np.load(GOLD / "ops_small.npz")


def test_rng_and_glorot(oracle, ops):
    oracle.init_rand_state(12345)
    assert oracle.get_rand_state() == tuple(int(v) for v in ops["rng_state_seed12345"])
    assert (oracle.rand(64) == ops["rng_first_64"]).all()
    oracle.init_rand_state(12345)
    assert same_bits(oracle.glorot(48, 16), ops["glorot_48x16"])
    assert same_bits(oracle.glorot(16, 7), ops["glorot_16x7"])


@pytest.mark.parametrize("dim", [7, 16, 41])
def test_graphsum(oracle, ops, dim):
    d = load_data(ops)
    out = oracle.graphsum(d.graph_indptr, d.graph_indices, ops[f"gs_in_{dim}"], dim)
    assert same_bits(out, ops[f"gs_fw_{dim}"])
    assert same_bits(out, ops[f"gs_bw_{dim}"])      # backward is the same loop (module.cpp:103-119)


def test_spmm_matmul(oracle, ops):
    d = load_data(ops)
    n = d.num_nodes
    args = (d.feature_indptr, d.feature_indices, d.feature_value)
    assert same_bits(oracle.spmm_fw(*args, ops["spmm_w"], n, 48, 16), ops["spmm_fw"])
    assert same_bits(oracle.spmm_bw(*args, ops["spmm_cgrad"], n, 48, 16), ops["spmm_bw"])
    assert same_bits(oracle.matmul_fw(ops["mm_a"], ops["mm_b"], n, 16, 7), ops["mm_fw"])
    a, b = oracle.matmul_bw(ops["mm_a"], ops["mm_b"], ops["mm_cgrad"], n, 16, 7)
    assert same_bits(a, ops["mm_bw_a"]) and same_bits(b, ops["mm_bw_b"])


def test_ce_relu_dropout_adam(oracle, ops):
    loss, shifted, grad = oracle.cross_entropy(ops["ce_logits"], ops["ce_truth"], 7, True)
    assert same_bits([loss], [ops["ce_loss"]]) and same_bits(shifted, ops["ce_shifted"]) and same_bits(grad, ops["ce_grad"])
    x, m, g = oracle.relu(ops["act_x"], ops["act_grad"], True)
    assert same_bits(x, ops["relu_x"]) and (m == ops["relu_mask"]).all() and same_bits(g, ops["relu_grad"])
    oracle.set_rand_state(*[int(v) for v in ops["drop_state"]])
    x, m, g = oracle.dropout(ops["act_x"], 0.5, ops["act_grad"], True, True)
    assert same_bits(x, ops["drop_x"]) and (m == ops["drop_mask"]).all() and same_bits(g, ops["drop_grad"])
    steps = [[ops["adam_g0"][i], ops["adam_g1"][i]] for i in range(len(ops["adam_g0"]))]
    r = oracle.adam([ops["adam_w0"], ops["adam_w1"]], steps, [1, 0], 0.01, 5e-4)
    assert same_bits(r[0], ops["adam_out0"]) and same_bits(r[1], ops["adam_out1"])


@pytest.mark.parametrize("tag", ["toy", "toy_dense"])
@pytest.mark.parametrize("drop", [0.0, 0.5])
def test_training_run(oracle, tag, drop):
    z = np.load(GOLD / f"train_{tag}.npz")
    d = load_data(z)
    key = f"p{int(drop * 10)}_"
    want = z[key + "epochs"]
    g = oracle.gcn(d, hidden_dim=16, dropout=drop, epochs=len(want), seed=int(z[key + "seed"][0]))
    assert same_bits(g.var(2), z[key + "w1_init"]) and same_bits(g.var(5), z[key + "w2_init"])
    for e in range(len(want)):
        row = (*g.train_epoch(), *g.eval(2))
        assert same_bits(row, want[e]), f"epoch {e}: {row} vs {want[e]}"
    assert same_bits(g.eval(3), z[key + "test"])
    assert same_bits(g.var(2), z[key + "w1_final"]) and same_bits(g.var(5), z[key + "w2_final"])
    assert same_bits(g.var(6), z[key + "logits_final"])
    g.close()


@pytest.mark.parametrize("name", ["toy", "quirks"])
def test_parser(oracle, name):
    want = np.load(GOLD / f"parser_{name}.npz")
    got = oracle.parse(GOLD / "parser_toy", name)
    assert got is not None
    for k in want.files:
        w = want[k]
        g = np.asarray(got[k])
        assert g.shape == w.shape, k
        assert (g.view(np.uint32) == w.view(np.uint32)).all() if w.dtype == np.float32 else (g == w).all(), k
