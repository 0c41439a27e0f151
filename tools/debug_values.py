"""debug: where do leaf values diverge from the compiled reference at 131072 rows?"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_baseline_sizes import _data, _engine, KEYS
cores = os.cpu_count()
def run(name, n, f, d, depth, grow, score, iters, T=None):
    T = T or cores
    seed = 4242 + n % 89 + f
    out = "/tmp/dbg_ref.npz"
    env = dict(os.environ, OMP_NUM_THREADS=str(T))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_fit_worker.py"), "fit", str(n), str(f), str(d), str(depth), grow, score, str(iters), str(seed), out], capture_output=True, text=True, env=env)
    assert "REF_WORKER_OK" in r.stdout, r.stderr[-500:]
    z = np.load(out)
    X, y = _data(n, f, d, seed)
    lrs = [(0.1, 0, 1)] if d == 1 else [(0.1, 0, d - 1), (0.01, d - 1, d)]
    m = _engine(f, d, depth, grow, score, n, lrs, T)
    loss = m.fit(X, None, y, iters, False, "MultiRMSE")
    e = m.get_ensemble_data()
    ti = z["fit_tree_indices"]
    same_struct = all(np.array_equal(np.asarray(z["fit_" + k]).astype(np.float64), np.asarray(e[k]).astype(np.float64)) for k in ("tree_indices", "depths", "feature_indices", "feature_values", "inequality_directions"))
    dv = np.abs(z["fit_values"].astype(np.float64) - e["values"])
    per_tree = [float(dv[ti[t]:(ti[t + 1] if t + 1 < len(ti) else dv.shape[0])].max()) for t in range(len(ti))]
    # exact leaf means of tree 0 from the reference's own bias: g = bias - y
    print(name, "n", n, "f", f, "T", T, "struct_equal", same_struct, "per-tree max|dv|", ["%.2e" % v for v in per_tree],
          "loss ref/ours %.7f %.7f" % (float(z["fit_loss"]), loss), "n_bad", int((dv > 1e-5).sum()), "of", dv.size, flush=True)
    bad = np.argwhere(dv > 1e-5)[:6]
    for b in bad:
        print("   leaf", b[0], "col", b[1], "ref", z["fit_values"][b[0], b[1]], "ours", e["values"][b[0], b[1]], "edge_w", z["fit_edge_weights"][b[0]].prod() * n)
    # bias check
    bias_np = y.astype(np.float64).mean(0)
    print("   bias ours", m.get_bias(), "float64 mean", bias_np, flush=True)
for T in (16, 4, 1):
    run("c2fam-f16", 131072, 16, 1, 6, "greedy", "L2", 1, T)
run("c2fam-f16-65k", 65536, 16, 1, 6, "greedy", "L2", 2)
run("c2fam-f16", 131072, 16, 1, 6, "greedy", "L2", 3)
run("j3fam-f16", 131072, 16, 1, 6, "oblivious", "cosine", 2)
run("c3fam-f16", 65536, 16, 2, 8, "oblivious", "cosine", 2)
run("c2fam", 131072, 128, 1, 6, "greedy", "L2", 1)
