#!/usr/bin/env python
"""tools/make_golden.py — regenerates tests/golden/*.npz from the UNMODIFIED reference
(oracle/_ref/libgcnref.so, built by `make -C oracle ref` from /root/reference).

The fixtures hold both the inputs and the reference's outputs, so the consumers
(tests/test_golden.py: the C restatement on CPU; tests/test_gpu_*.py: the CUDA path) need neither
/root/reference nor numpy's RNG to be stable.  Run from the repo root:  python tools/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.checker import Ref  # noqa: E402
from tests.util import make_dataset, write_text_dataset  # noqa: E402

OUT = ROOT / "tests" / "golden"


def data_dict(d, prefix="d_"):
    return {prefix + k: getattr(d, k) for k in ("graph_indptr", "graph_indices", "feature_indptr", "feature_indices",
                                                "feature_value", "label", "split")} | {
        prefix + "dims": np.array([d.num_nodes, d.input_dim, d.output_dim], np.int32)}


def ops(ref: Ref):
    rng = np.random.default_rng(2024)
    out = {}
    d = make_dataset(n=180, f=48, c=7, n_undirected=700, nnz_per_row=6, seed=5, isolated=6, hub=(4, 120))
    out |= data_dict(d)
    n = d.num_nodes
    ref.init_rand_state(12345)
    out["rng_state_seed12345"] = np.array(ref.get_rand_state(), np.uint64)
    out["rng_first_64"] = ref.rand(64)
    ref.init_rand_state(12345)
    out["glorot_48x16"] = ref.glorot(48, 16)
    out["glorot_16x7"] = ref.glorot(16, 7)
    for dim in (7, 16, 41):
        x = rng.standard_normal(n * dim).astype(np.float32)
        out[f"gs_in_{dim}"] = x
        out[f"gs_fw_{dim}"] = ref.graphsum(d.graph_indptr, d.graph_indices, x, dim)
        out[f"gs_bw_{dim}"] = ref.graphsum(d.graph_indptr, d.graph_indices, x, dim, backward=True)
    w = rng.standard_normal(48 * 16).astype(np.float32)
    g = rng.standard_normal(n * 16).astype(np.float32)
    out["spmm_w"], out["spmm_cgrad"] = w, g
    out["spmm_fw"] = ref.spmm_fw(d.feature_indptr, d.feature_indices, d.feature_value, w, n, 48, 16)
    out["spmm_bw"] = ref.spmm_bw(d.feature_indptr, d.feature_indices, d.feature_value, g, n, 48, 16)
    a, b, cg = (rng.standard_normal(s).astype(np.float32) for s in (n * 16, 16 * 7, n * 7))
    out["mm_a"], out["mm_b"], out["mm_cgrad"] = a, b, cg
    out["mm_fw"] = ref.matmul_fw(a, b, n, 16, 7)
    out["mm_bw_a"], out["mm_bw_b"] = ref.matmul_bw(a, b, cg, n, 16, 7)
    logits = (rng.standard_normal(n * 7) * 2).astype(np.float32)
    truth = rng.integers(-1, 7, n).astype(np.int32)
    out["ce_logits"], out["ce_truth"] = logits, truth
    loss, shifted, grad = ref.cross_entropy(logits, truth, 7, True)
    out["ce_loss"], out["ce_shifted"], out["ce_grad"] = np.float32(loss), shifted, grad
    x = rng.standard_normal(n * 16).astype(np.float32)
    gr = rng.standard_normal(n * 16).astype(np.float32)
    out["act_x"], out["act_grad"] = x, gr
    xr, mr, grr = ref.relu(x, gr, True)
    out["relu_x"], out["relu_mask"], out["relu_grad"] = xr, mr, grr
    ref.set_rand_state(88172645463325252, 1181783497276652981)
    xd, md, gd = ref.dropout(x, 0.5, gr, True, True)
    out["drop_state"] = np.array([88172645463325252, 1181783497276652981], np.uint64)
    out["drop_x"], out["drop_mask"], out["drop_grad"] = xd, md, gd
    datas = [rng.standard_normal(300).astype(np.float32), rng.standard_normal(40).astype(np.float32)]
    steps = [[rng.standard_normal(300).astype(np.float32) * 1e-2, rng.standard_normal(40).astype(np.float32)] for _ in range(10)]
    res = ref.adam(datas, steps, [1, 0], 0.01, 5e-4)
    out["adam_w0"], out["adam_w1"] = datas
    out["adam_g0"] = np.stack([s[0] for s in steps])
    out["adam_g1"] = np.stack([s[1] for s in steps])
    out["adam_out0"], out["adam_out1"] = res
    np.savez_compressed(OUT / "ops_small.npz", **out)


def training(ref: Ref):
    for tag, kw in (("toy", dict(n=400, f=96, c=6, n_undirected=1500, nnz_per_row=9, seed=31, isolated=5)),
                    ("toy_dense", dict(n=256, f=32, c=5, n_undirected=2000, seed=32, dense=True, alpha=1.5))):
        d = make_dataset(**kw)
        out = data_dict(d)
        for drop in (0.0, 0.5):
            epochs, seed = 12, 2019
            g = ref.gcn(d, hidden_dim=16, dropout=drop, epochs=epochs, seed=seed)
            key = f"p{int(drop * 10)}_"
            out[key + "w1_init"], out[key + "w2_init"] = g.var(2), g.var(5)
            rows = []
            for _ in range(epochs):
                tl, ta = g.train_epoch()
                vl, va = g.eval(2)
                rows.append((tl, ta, vl, va))
            out[key + "epochs"] = np.array(rows, np.float32)
            out[key + "test"] = np.array(g.eval(3), np.float32)
            out[key + "w1_final"], out[key + "w2_final"] = g.var(2), g.var(5)
            out[key + "logits_final"] = g.var(6)
            out[key + "seed"] = np.array([seed], np.int64)
            g.close()
        np.savez_compressed(OUT / f"train_{tag}.npz", **out)


def parser(ref: Ref):
    d = make_dataset(n=50, f=24, c=4, n_undirected=120, nnz_per_row=4, seed=77, isolated=3, empty_rows=2)
    root = OUT / "parser_toy"
    write_text_dataset(root, "toy", d)
    # the accepted-input quirks of SURVEY Appendix B
    q = root / "data"
    (q / "quirks.graph").write_text("1 2\n0 0 x 5\n\n 2   1 \n0 1")
    (q / "quirks.split").write_text("1\n2\n3\n0\n")
    (q / "quirks.svmlight").write_text("0 0:1.5 3:2\n2 1:0.25\n\n1 2:1e-3 2:7\n")
    for name in ("toy", "quirks"):
        p = ref.parse(root, name)
        np.savez_compressed(OUT / f"parser_{name}.npz", **{k: np.asarray(v) for k, v in p.items()})


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    r = Ref()
    ops(r)
    training(r)
    parser(r)
    print("golden fixtures written to", OUT)
