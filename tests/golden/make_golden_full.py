"""Full-size known-answer vectors for the BASELINE configs, generated from the REFERENCE ITSELF (oracle/_ref).

    OMP_NUM_THREADS=8 python tests/golden/make_golden_full.py c2 [iterations]

Runs the reference's own CPU path at the FULL size of a BASELINE config (minutes to hours per boosting iteration on 8
cores) -- GBRL.fit (shuffle=False, MultiRMSE, batch_size = N) for the "fit" fixtures, a GBRL.step boosting loop driven
by the reference's own predictions for the "step" fixtures -- and stores ONLY what is needed to pin parity:
the seed / shape / hyper-parameters, the thread count the reference ran with (its mean / std / loss reductions
are thread-partitioned, so the engine has to emulate the same partition: ref_threads), the resulting ensemble
arrays (a few KB), the loss, the bias and a prefix of the predictions.  The inputs are regenerated from the seed
by the test (numpy Generator streams are stable across machines).

Why "step" fixtures: the reference's MultiRMSE::get_loss_and_gradients (loss.cpp:42-57) and calculate_std_and_center
(math_ops.cpp:459-487) keep their per-element temporary (`grad_value`, `value`) in a variable declared OUTSIDE the
`omp parallel` region, i.e. shared by all threads: a data race that, depending on timing, corrupts a few gradients of a
fit() run with many threads (observed: 7 of 64 leaf values of a 131072 x 128 fit off by ~3e-4 with 16 threads, leaf values
that are no longer the means of their samples; none with 8 threads here).  The step() path has no such race.  A "step"
fixture drives the boosting loop with gradients g_t = predict_t(X) - y computed outside the reference, the test rebuilds
the same stream from the stored ensemble (tests/helpers.py numpy_predict, asserted bit-identical to the reference's
predictions here), and every fixture is checked for self-consistency (leaf values == means of the leaf's gradients).
"""
import os
import sys
import time

T = int(os.environ.get("OMP_NUM_THREADS", "0"))
assert T >= 1, "set OMP_NUM_THREADS explicitly: the thread count is part of the fixture"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from oracle.oracle import load_reference, make_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("tree_indices", "depths", "feature_indices", "feature_values", "inequality_directions", "edge_weights", "values")

FULL = {
    # BASELINE.json configs at their true sizes (SURVEY 8d hyper-parameters)
    "c1": dict(n=10_000, f=16, d=1, depth=4, grow="oblivious", score="cosine", lrs=[(0.1, 0, 1)], iters=10),
    "c1_l2": dict(n=10_000, f=16, d=1, depth=4, grow="oblivious", score="L2", lrs=[(0.1, 0, 1)], iters=10),
    "c2": dict(n=1_000_000, f=128, d=1, depth=6, grow="greedy", score="L2", lrs=[(0.1, 0, 1)], iters=2),
    "c3": dict(n=4_000_000, f=64, d=2, depth=8, grow="oblivious", score="cosine", lrs=[(0.1, 0, 1), (0.01, 1, 2)], iters=1, mode="step"),
    "c5": dict(n=8_000_000, f=256, d=1, depth=6, grow="greedy", score="L2", lrs=[(0.1, 0, 1)], iters=1, mode="step"),
    "c2s": dict(n=1_000_000, f=128, d=1, depth=6, grow="greedy", score="L2", lrs=[(0.1, 0, 1)], iters=2, mode="step"),
    # small step-mode fixtures (they exercise numpy_predict and the step replay of the test on both tree layouts)
    "s_greedy": dict(n=20_000, f=16, d=2, depth=5, grow="greedy", score="cosine", lrs=[(0.1, 0, 1), (0.01, 1, 2)], iters=3, mode="step"),
    "s_obl": dict(n=20_000, f=16, d=2, depth=5, grow="oblivious", score="L2", lrs=[(0.1, 0, 1), (0.01, 1, 2)], iters=3, mode="step"),
    # north_star's target workload: 1M x 128 oblivious fit
    "j3": dict(n=1_000_000, f=128, d=1, depth=6, grow="oblivious", score="cosine", lrs=[(0.1, 0, 1)], iters=1),
}


def data(n, f, d, seed):
    """Same generator as bench.py synth_numpy (float32 streams)."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, f), dtype=np.float32)
    W = rng.standard_normal((f, d), dtype=np.float32)
    y = (np.tanh(X @ W / np.sqrt(f)) + 0.1 * rng.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
    return X, y


def main():
    name = sys.argv[1]
    c = dict(FULL[name])
    if len(sys.argv) > 2:
        c["iters"] = int(sys.argv[2])
    ref = load_reference()
    assert ref is not None, "oracle/_ref is not built (make -C oracle ref)"
    n, f, d = c["n"], c["f"], c["d"]
    seed = 1234 + n % 97 + f
    X, y = data(n, f, d, seed)
    m = make_reference(ref, input_dim=f, output_dim=d, max_depth=c["depth"], n_bins=256, par_th=10,
                       split_score_func=c["score"], generator_type="quantile", batch_size=n, grow_policy=c["grow"],
                       lrs=c["lrs"])
    mode = c.get("mode", "fit")
    t0 = time.time()
    if mode == "fit":
        loss = m.fit(X, None, y, c["iters"], False, "MultiRMSE")
        bias = np.array(m.get_bias(), copy=True)
        preds = []
    else:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        bias = y.astype(np.float64).mean(0).astype(np.float32)
        m.set_bias(bias)
        preds, loss = [], 0.0
        for it in range(c["iters"]):
            p = np.array(m.predict(X, None), copy=True).reshape(n, d)
            preds.append(p)
            m.step(X, None, (p - y).astype(np.float32))
    dt = time.time() - t0
    e = m.get_ensemble_data()
    if mode == "step":
        from helpers import numpy_predict
        for it, p in enumerate(preds):      # the test rebuilds these from the stored ensemble: they must be the reference's bits
            q = numpy_predict(e, X, bias, c["lrs"], it, c["grow"] == "oblivious")
            assert np.array_equal(q.view(np.uint32), p.view(np.uint32)), "numpy_predict differs from the reference at iteration %d" % it
    pred = np.array(m.predict(X[:65536], None), copy=True).reshape(-1, d)
    out = {"cfg": np.array([n, f, d, c["depth"], 256, c["iters"], n, seed, T], np.int64), "score": c["score"], "grow": c["grow"],
           "gen": "quantile", "lrs": np.array(c["lrs"], np.float32), "fit_loss": np.float32(loss), "mode": mode,
           "fit_bias": bias, "fit_pred_head": pred, "seconds": np.float64(dt)}
    for k in KEYS:
        out["fit_%s" % k] = np.array(e[k], copy=True)
    np.savez_compressed(os.path.join(HERE, "full_%s.npz" % name), **out)
    print("wrote full_%s: %d iterations in %.1f s (%d threads), %d leaves, loss %.6f" % (
        name, c["iters"], dt, T, out["fit_values"].shape[0], loss), flush=True)
    os._exit(0)   # the reference's get_ensemble_data capsules double-free at teardown (see make_golden.py)


if __name__ == "__main__":
    main()
