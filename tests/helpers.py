"""Shared helpers of the parity tests: synthetic data, engine adaptors, ensemble comparison."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INT_KEYS = ("tree_indices", "depths", "feature_indices", "inequality_directions")
THR_KEYS = ("feature_values",)
FLOAT_KEYS = ("values", "edge_weights")
TOL = 1e-5   # north_star: "within 1e-5 on fp32 leaf values/predictions"


def synth(n, f, d, seed):
    """X ~ N(0,1), targets = tanh(XW/sqrt(F)) + 0.1 eps  (SURVEY.md 8d)."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, f)).astype(np.float32)
    W = rng.standard_normal((f, d)).astype(np.float32)
    y = (np.tanh(X @ W / np.sqrt(f)) + 0.1 * rng.standard_normal((n, d))).astype(np.float32)
    return X, y


def default_lrs(d):
    return [(0.1, 0, d)] if d == 1 else [(0.1, 0, d - 1), (0.05, d - 1, d)]


def configure(model, f, d, lrs=None, fw=None, bias=None):
    model.set_bias(np.zeros(d, np.float32) if bias is None else np.asarray(bias, np.float32))
    model.set_feature_weights(np.ones(f, np.float32) if fw is None else np.asarray(fw, np.float32))
    model.set_feature_mapping(np.arange(f, dtype=np.int32), np.ones(f, dtype=bool))
    for (lr, a, b) in (lrs or default_lrs(d)):
        model.set_optimizer("SGD", "const", float(lr), int(a), int(b))
    return model


def make_oracle(ref_threads=1, **kw):
    from oracle.oracle import Oracle
    return Oracle(ref_threads=ref_threads, **kw)


def make_gpu(ref_threads=1, **kw):
    from gbrl_b200 import GBRL
    kw = dict(kw)
    kw.setdefault("policy_dim", kw["output_dim"])
    return GBRL(device="cuda", ref_threads=ref_threads, **kw)


class OracleAdaptor:
    def __init__(self, o):
        self.o = o

    def step(self, X, g):
        self.o.step(X, g)

    def predict(self, X, start=0, stop=0):
        return np.asarray(self.o.predict(X, start, stop))

    def fit(self, X, y, iters):
        return self.o.fit(X, y, iters)

    def ensemble(self):
        return self.o.get_ensemble_data()


class GpuAdaptor:
    def __init__(self, m):
        self.m = m

    def step(self, X, g):
        self.m.step(X, None, g)

    def predict(self, X, start=0, stop=0):
        return self.m.predict_numpy(X, start, stop)

    def fit(self, X, y, iters):
        return self.m.fit(X, None, y, iters, False, "MultiRMSE")

    def ensemble(self):
        return self.m.get_ensemble_data()


def compare_ensembles(a, b, what="", tol=TOL, n_trees=None):
    """a = expected (reference / oracle), b = candidate.  Bit-exact on integers, thresholds and directions;
    |.| <= tol on leaf values and edge weights."""
    for k in INT_KEYS + THR_KEYS:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape, "%s %s shape %s vs %s" % (what, k, x.shape, y.shape)
        if k in THR_KEYS:
            same = np.array_equal(x.view(np.uint32), y.view(np.uint32)) or np.array_equal(x, y)
        else:
            same = np.array_equal(x.astype(np.int64), y.astype(np.int64))
        assert same, "%s %s differs at %s" % (what, k, np.argwhere(x != y)[:4].tolist())
    for k in FLOAT_KEYS:
        x = np.asarray(a[k], np.float64)
        y = np.asarray(b[k], np.float64).reshape(x.shape)
        err = np.abs(x - y).max() if x.size else 0.0
        assert err <= tol, "%s %s max abs err %.3e > %.1e" % (what, k, err, tol)


def boosting_loop(engines, X, y, iters, check_each=True):
    """Drive several engines with the SAME gradient stream (computed from the first engine's predictions,
    i.e. the expected one) so that a mismatch cannot hide behind diverging gradients."""
    n = X.shape[0]
    d = y.shape[1]
    for it in range(iters):
        preds = [np.asarray(e.predict(X)).reshape(n, d) for e in engines]
        for p in preds[1:]:
            err = np.abs(p.astype(np.float64) - preds[0]).max()
            assert err <= TOL, "iteration %d: prediction err %.3e" % (it, err)
        g = (preds[0] - y).astype(np.float32)
        for e in engines:
            e.step(X, g)
        if check_each:
            ens = [e.ensemble() for e in engines]
            for other in ens[1:]:
                compare_ensembles(ens[0], other, "iteration %d" % it)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    n, f, d, depth, bins, iters, batch = [int(v) for v in z["cfg"]]
    cfg = dict(input_dim=f, output_dim=d, max_depth=depth, n_bins=bins, par_th=10, split_score_func=str(z["score"]),
               generator_type=str(z["gen"]), batch_size=batch, grow_policy=str(z["grow"]))
    lrs = [(float(a), int(b), int(c)) for a, b, c in z["lrs"]]
    return z, cfg, lrs, iters, str(z["mode"])


def run_golden(name, factory, adaptor):
    """Replays a fixture generated from the reference (tests/golden/make_golden.py, OMP_NUM_THREADS=1)."""
    z, cfg, lrs, iters, mode = load_golden(name)
    X, y = z["X"], z["y"]
    n, d = y.shape
    m = factory(ref_threads=1, **cfg)
    configure(m, cfg["input_dim"], d, lrs=lrs, fw=z["fw"])
    e = adaptor(m)
    if mode == "step":
        for it in range(iters):
            p = np.asarray(e.predict(X)).reshape(n, d)
            if it > 0:
                assert np.abs(p.astype(np.float64) - z["it%d_pred" % (it - 1)]).max() <= TOL
            # gradients from the REFERENCE's predictions, as stored
            pref = z["it%d_pred" % (it - 1)] if it > 0 else np.zeros((n, d), np.float32)
            g = (pref - y).astype(np.float32)
            e.step(X, g)
        exp = {k: z["final_" + k] for k in INT_KEYS + THR_KEYS + FLOAT_KEYS}
        compare_ensembles(exp, e.ensemble(), name)
        assert np.abs(np.asarray(e.predict(X)).reshape(n, d).astype(np.float64) - z["it%d_pred" % (iters - 1)]).max() <= TOL
    else:
        loss = e.fit(X, y, iters)
        exp = {k: z["fit_" + k] for k in INT_KEYS + THR_KEYS + FLOAT_KEYS}
        compare_ensembles(exp, e.ensemble(), name)
        assert abs(loss - float(z["fit_loss"])) <= 1e-5 * max(1.0, abs(float(z["fit_loss"])))
        assert np.abs(np.asarray(e.predict(X)).reshape(n, d).astype(np.float64) - z["fit_pred"]).max() <= TOL
    return e


def numpy_predict(e, X, bias, lrs, n_trees, oblivious):
    """The reference's sample-parallel predict restated in numpy float32 (predictor.cpp:167-265, optimizer.cpp:110-118):
    theta = bias, then for every tree in ascending order theta[d] = fl(theta[d] - fl(lr_d * value[leaf, d])).  Bit-identical to
    the reference's predictions (asserted when the full-size fixtures are generated), so a test can rebuild the exact
    gradient stream of a boosting run from the stored ensemble alone."""
    X = np.asarray(X, np.float32)
    n, D = X.shape[0], len(bias)
    lr_d = np.zeros(D, np.float32)
    for (lr, a, b) in lrs:
        lr_d[int(a):int(b)] = np.float32(lr)
    out = np.tile(np.asarray(bias, np.float32), (n, 1))
    ti, nl = np.asarray(e["tree_indices"]), np.asarray(e["values"]).shape[0]
    vals = np.asarray(e["values"], np.float32).reshape(nl, D)
    fi, fv, iq, dp = (np.asarray(e[k]) for k in ("feature_indices", "feature_values", "inequality_directions", "depths"))
    for t in range(n_trees):
        l0, l1 = int(ti[t]), (int(ti[t + 1]) if t + 1 < len(ti) else nl)
        if oblivious:
            dep = int(dp[t])                               # depth 0: leaf 0 is applied (predictor.cpp:244-258); its value is 0
            li = np.zeros(n, np.int64)
            for k in range(dep):
                li |= (X[:, int(fi[t, k])] > fv[t, k]).astype(np.int64) << (dep - 1 - k)
            leaf = l0 + li
        else:
            leaf = np.full(n, -1, np.int64)
            for L in range(l0, l1):
                dep = int(dp[L])
                if dep == 0:
                    continue
                ok = np.ones(n, bool)
                for k in range(dep):
                    ok &= (X[:, int(fi[L, k])] > fv[L, k]) == bool(iq[L, k])
                leaf[ok] = L
        hit = leaf >= 0
        upd = (lr_d[None, :] * vals[np.where(hit, leaf, 0)]).astype(np.float32)
        out = np.where(hit[:, None], (out - upd).astype(np.float32), out)
    return out
