import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_baseline_sizes import _data, _engine, KEYS
n, f, d, depth = 131072, 128, 1, 6
seed = 4242 + n % 89 + f
X, y = _data(n, f, d, seed)
m = _engine(f, d, depth, "greedy", "L2", n, [(0.1, 0, 1)], 16)
loss = m.fit(X, None, y, 1, False, "MultiRMSE")
e = m.get_ensemble_data()
bias = m.get_bias()
g = (bias[None, :] - y).astype(np.float32)
nl = e["values"].shape[0]
member = np.full(n, -1)
for leaf in range(nl):
    dep = int(e["depths"][leaf]); ok = np.ones(n, bool)
    for k in range(dep):
        ok &= (X[:, e["feature_indices"][leaf, k]] > e["feature_values"][leaf, k]) == bool(e["inequality_directions"][leaf, k])
    member[ok] = leaf
means = np.array([g[member == l].astype(np.float64).mean() if (member == l).any() else 0 for l in range(nl)])
cnts = np.array([(member == l).sum() for l in range(nl)])
dv = np.abs(means - e["values"][:, 0])
print("featmajor env", os.environ.get("GBRL_B200_FEATMAJOR"), "max |host mean - engine value|", dv.max(), "bad leaves", np.argwhere(dv > 1e-5).ravel().tolist())
for l in np.argwhere(dv > 1e-5).ravel()[:8]:
    w = np.prod(e["edge_weights"][l, :int(e["depths"][l])].astype(np.float64)) * n
    print("  leaf", l, "host cnt", cnts[l], "engine cnt(edge w)", w, "host mean", means[l], "engine", e["values"][l, 0], "path", [(int(e["feature_indices"][l,k]), float(e["feature_values"][l,k]), bool(e["inequality_directions"][l,k])) for k in range(int(e["depths"][l]))])
# the reference's own ensemble on the same data: are ITS values the host means?
out = "/tmp/dbg_ref2.npz"
env = dict(os.environ, OMP_NUM_THREADS="16")
r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_fit_worker.py"), "fit", str(n), str(f), str(d), str(depth), "greedy", "L2", "1", str(seed), out], capture_output=True, text=True, env=env)
z = np.load(out)
rv = z["fit_values"][:, 0]
print("reference vs host means: max", np.abs(means - rv).max(), "| engine vs reference: max", np.abs(e["values"][:, 0] - rv).max())
rb = (rv - means)
print("ref - host mean per bad leaf", [(int(l), float(rb[l]), int(cnts[l])) for l in np.argwhere(np.abs(rb) > 1e-5).ravel()[:8]])
