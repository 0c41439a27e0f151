"""Runs the REFERENCE's own CPU path (oracle/_ref, compiled from the reference sources) on a seeded synthetic matrix with
all the host threads OMP_NUM_THREADS grants, and dumps the ensemble.  Worker of tests/test_gpu_baseline_sizes.py: a
separate process, because the reference's thread count (which fixes the partition of its float reductions) is latched
from the environment when libgomp loads.  TEST INFRASTRUCTURE ONLY.

    OMP_NUM_THREADS=16 python tests/ref_fit_worker.py fit  n f d depth grow score iters seed out.npz
    OMP_NUM_THREADS=16 python tests/ref_fit_worker.py step n f d depth grow score iters seed out.npz
    OMP_NUM_THREADS=16 python tests/ref_fit_worker.py load model.gbrl_model obs.npy out.npy
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from oracle.oracle import load_reference, make_reference  # noqa: E402

KEYS = ("tree_indices", "depths", "feature_indices", "feature_values", "inequality_directions", "edge_weights", "values")


def data(n, f, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, f), dtype=np.float32)
    W = rng.standard_normal((f, d), dtype=np.float32)
    y = (np.tanh(X @ W / np.sqrt(f)) + 0.1 * rng.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
    return X, y


def lrs_of(d):
    return [(0.1, 0, 1)] if d == 1 else [(0.1, 0, d - 1), (0.01, d - 1, d)]


def main():
    ref = load_reference()
    assert ref is not None, "oracle/_ref is not built"
    if sys.argv[1] == "load":
        m = ref.GBRL.load(sys.argv[2])
        X = np.load(sys.argv[3])
        t = time.time()
        p = np.array(m.predict(X, None), copy=True)
        np.save(sys.argv[4], p)
        print("REF_WORKER_OK predict %.2f s" % (time.time() - t), flush=True)
        os._exit(0)
    mode = sys.argv[1]
    n, f, d, depth = [int(v) for v in sys.argv[2:6]]
    grow, score, iters, seed, out = sys.argv[6], sys.argv[7], int(sys.argv[8]), int(sys.argv[9]), sys.argv[10]
    X, y = data(n, f, d, seed)
    m = make_reference(ref, input_dim=f, output_dim=d, max_depth=depth, n_bins=256, par_th=10, split_score_func=score,
                       generator_type="quantile", batch_size=n, grow_policy=grow, lrs=lrs_of(d))
    t = time.time()
    if mode == "fit":
        loss = m.fit(X, None, y, iters, False, "MultiRMSE")
        bias = np.array(m.get_bias(), copy=True)
    else:
        # GBRL.step boosting loop, gradients computed OUTSIDE the reference: its fit() path races on a shared temporary in
        # MultiRMSE::get_loss_and_gradients (loss.cpp:42-57) when it runs on many threads (tests/golden/make_golden_full.py)
        bias = y.astype(np.float64).mean(0).astype(np.float32)
        m.set_bias(bias)
        loss = 0.0
        for it in range(iters):
            p = np.array(m.predict(X, None), copy=True).reshape(n, d)
            m.step(X, None, (p - y).astype(np.float32))
    dt = time.time() - t
    e = m.get_ensemble_data()
    o = {"fit_%s" % k: np.array(e[k], copy=True) for k in KEYS}
    o["fit_loss"] = np.float32(loss)
    o["fit_bias"] = bias
    o["fit_pred_head"] = np.array(m.predict(X[:8192], None), copy=True).reshape(-1, d)
    o["seconds"] = np.float64(dt)
    np.savez(out, **o)
    print("REF_WORKER_OK fit %d iterations in %.1f s on %s threads" % (iters, dt, os.environ.get("OMP_NUM_THREADS")), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
