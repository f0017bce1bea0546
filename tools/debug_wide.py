"""Debug aid: one training epoch of the wide plan against the reference, layer by layer."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cuda_gcn_b200 import abi, host_api
from oracle.checker import GraphData, best_checker
abi.require_device(0)
d = host_api.Data.synth("products", 0.002)
a = d.arrays()
gd = GraphData(a["graph_indptr"], a["graph_indices"], a["feature_indptr"], a["feature_indices"], a["feature_value"], a["label"], a["split"],
               input_dim=d.params.input_dim, output_dim=d.params.output_dim)
chk = best_checker()
for hidden, p in ((256, 0.0), (256, 0.5), (64, 0.5)):
    ref = chk.gcn(gd, hidden_dim=hidden, dropout=p, epochs=2, seed=7)
    eng = host_api.Engine(d, hidden_dim=hidden, dropout=p, epochs=2, seed=7, plan=host_api.PLAN_FUSED)
    w, g = ref.train_epoch(), eng.train_epoch()
    print(f"hidden {hidden} p {p}: train loss ref {w[0]:.6f} ours {g[0]:.6f}  acc {w[1]:.5f} {g[1]:.5f}")
    h_ref, h_eng = ref.var(3), eng.var(3)
    print("  H1: max|ref| %.4g max diff %.4g  zero pattern agreement %.6f" % (np.abs(h_ref).max(), np.abs(h_ref - h_eng).max(), ((h_ref == 0) == (h_eng == 0)).mean()))
    C = d.params.output_dim
    tr = (gd.split == 1) & (gd.label >= 0)
    lr, le = ref.var(6).reshape(-1, C)[tr], eng.var(6).reshape(-1, C)[tr]
    lr = lr - lr.max(1, keepdims=True); le = le - le.max(1, keepdims=True)
    print("  logits(train rows, shifted): max|ref| %.4g max diff %.4g" % (np.abs(lr).max(), np.abs(lr - le).max()))
    for idx in (2, 5):
        x, y = ref.var(idx, True), eng.var(idx, True)
        print(f"  grad W{1 if idx == 2 else 2}: max|ref| {np.abs(x).max():.4g} max diff {np.abs(x - y).max():.4g}")
    gr, ge = ref.var(3, True), eng.var(3, True)
    print("  dH1(masked): max|ref| %.4g max diff %.4g" % (np.abs(gr).max(), np.abs(gr - ge).max()))
    w, g = ref.eval(2), eng.eval(2)
    print(f"  eval: loss ref {w[0]:.6f} ours {g[0]:.6f}")
    ref.close(); eng.close()
