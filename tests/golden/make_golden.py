"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref, compiled from /root/reference by
oracle/Makefile).  Run in the build container only:

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py

Each fixture stores the inputs (X, y), the hyper-parameters and, after every boosting iteration driven
through gbrl_cpp.GBRL.step (or one GBRL.fit call), the reference's ensemble arrays and predictions.
OMP_NUM_THREADS=1 pins the reference's thread-partitioned float reductions (ref_threads=1).
"""
import os
import sys

assert os.environ.get("OMP_NUM_THREADS") == "1", "run with OMP_NUM_THREADS=1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from oracle.oracle import load_reference, make_reference  # noqa: E402

ref = load_reference()
assert ref is not None, "oracle/_ref is not built (make -C oracle ref)"
HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("tree_indices", "depths", "feature_indices", "feature_values", "inequality_directions", "edge_weights", "values")


def data(n, f, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, f)).astype(np.float32)
    W = rng.standard_normal((f, d)).astype(np.float32)
    y = (np.tanh(X @ W / np.sqrt(f)) + 0.1 * rng.standard_normal((n, d))).astype(np.float32)
    return X, y


CASES = [
    # name, n, f, d, depth, bins, score, grow, gen, iters, mode, batch
    ("greedy_l2", 1500, 7, 1, 5, 64, "L2", "greedy", "quantile", 4, "step", 0),
    ("greedy_cos_ac", 1200, 6, 3, 4, 32, "cosine", "greedy", "quantile", 4, "step", 0),
    ("obl_cos_ac", 1400, 5, 2, 4, 48, "cosine", "oblivious", "quantile", 4, "step", 0),
    ("obl_l2_uniform", 1000, 9, 1, 3, 40, "L2", "oblivious", "uniform", 3, "step", 0),
    ("fit_greedy_l2_mb", 1300, 6, 2, 3, 32, "L2", "greedy", "uniform", 5, "fit", 500),
    ("fit_obl_cos", 900, 4, 1, 4, 64, "cosine", "oblivious", "quantile", 4, "fit", 900),
]

KEEP = []
for (name, n, f, d, depth, bins, score, grow, gen, iters, mode, batch) in CASES:
    X, y = data(n, f, d, len(name) * 7 + n)
    lrs = [(0.1, 0, d)] if d == 1 else [(0.1, 0, d - 1), (0.05, d - 1, d)]
    fw = (1.0 + 0.1 * np.arange(f)).astype(np.float32)
    m = make_reference(ref, input_dim=f, output_dim=d, max_depth=depth, n_bins=bins, par_th=10, split_score_func=score,
                       generator_type=gen, batch_size=batch if batch else n, grow_policy=grow, lrs=lrs, feature_weights=fw)
    out = {"X": X, "y": y, "fw": fw, "lrs": np.array(lrs, np.float32),
           "cfg": np.array([n, f, d, depth, bins, iters, batch if batch else n], np.int64),
           "score": score, "grow": grow, "gen": gen, "mode": mode}
    if mode == "step":
        for it in range(iters):
            p = np.array(m.predict(X, None), copy=True).reshape(n, d)
            g = (p - y).astype(np.float32)
            m.step(X, None, g)
            out["it%d_pred" % it] = np.array(m.predict(X, None), copy=True).reshape(n, d)
        # NOTE: the reference's get_ensemble_data() may only be called once per model (its capsules free
        # buffers a second call would free again), so the ensemble is dumped once, after the last step;
        # trees are append-only, so the state after iteration i is a prefix of these arrays.
        e = m.get_ensemble_data()
        KEEP.append(e)
        for k in KEYS:
            out["final_%s" % k] = np.array(e[k], copy=True)
    else:
        loss = m.fit(X, None, y, iters, False, "MultiRMSE")
        e = m.get_ensemble_data()
        KEEP.append(e)
        for k in KEYS:
            out["fit_%s" % k] = np.array(e[k], copy=True)
        out["fit_loss"] = np.float32(loss)
        out["fit_bias"] = np.array(m.get_bias(), copy=True)
        out["fit_pred"] = np.array(m.predict(X, None), copy=True).reshape(n, d)
    for k, v in out.items():
        v = np.asarray(v)
        assert v.size < 10**7, (k, v.shape, v.dtype)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:3]})

sys.stdout.flush()
os._exit(0)   # skip the reference module's teardown (see NOTE above)
