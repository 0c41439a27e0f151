"""GPU suite: BASELINE.json's full sizes, checked through size-independent properties (the oracle needs minutes to
hours per iteration there): histogram conservation, partition consistency, determinism of the integer path,
predict == sum over trees, multi-GPU == single GPU."""
import numpy as np
import pytest

from helpers import GpuAdaptor, configure, make_gpu, synth, TOL

pytestmark = pytest.mark.gpu


def _check_tree_invariants(e, X, grads, md, oblivious, lr=0.1):
    """Leaf assignment recomputed on the host from the emitted paths must partition the samples, the leaf values
    must be the per-leaf means of the raw gradients, and edge weights must be the child/parent count ratios."""
    n, D = grads.shape
    nl = e["values"].shape[0]
    counts = np.zeros(nl, np.int64)
    member = np.full(n, -1, np.int64)
    for leaf in range(nl):
        dep = int(e["depths"][0] if oblivious else e["depths"][leaf])
        row = 0 if oblivious else leaf
        ok = np.ones(n, bool)
        for k in range(dep):
            fidx, thr = int(e["feature_indices"][row, k]), e["feature_values"][row, k]
            ok &= (X[:, fidx] > thr) == bool(e["inequality_directions"][leaf, k])
        assert np.all(member[ok] == -1), "leaves overlap"
        member[ok] = leaf
        counts[leaf] = ok.sum()
        if ok.any():
            mean = grads[ok].astype(np.float64).mean(0)
            assert np.abs(mean - e["values"][leaf]).max() <= 2e-5 * max(1.0, np.abs(mean).max())
        w = np.prod(e["edge_weights"][leaf, :dep].astype(np.float64)) if dep else 1.0
        assert abs(w - counts[leaf] / n) <= 1e-4
    assert np.all(member >= 0), "leaves do not cover all samples"
    return member


@pytest.mark.parametrize("grow,score,n,f,d,depth", [
    ("greedy", "L2", 1_000_000, 128, 1, 6),       # BASELINE config 2
    ("oblivious", "cosine", 1_000_000, 64, 2, 8), # BASELINE config 3 at N/4 (full N in bench.py)
])
def test_full_size_tree_invariants(grow, score, n, f, d, depth):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, f), dtype=np.float32)
    W = rng.standard_normal((f, d), dtype=np.float32)
    y = (np.tanh(X @ W / np.sqrt(f)) + 0.1 * rng.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
    kw = dict(input_dim=f, output_dim=d, max_depth=depth, n_bins=256, split_score_func=score, generator_type="quantile",
              batch_size=n, grow_policy=grow, ref_threads=64)
    m = configure(make_gpu(**kw), f, d)
    grads = (0.0 - y).astype(np.float32)
    m.step(X, None, grads)
    e = m.get_ensemble_data()
    st = m.get_stats()
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/noise_stats.txt", "a") as fh:
        fh.write("full-size %s %s n=%d | replay_nodes %d / %d evaluated, items %d, max_noise_ratio %.3f\n" % (
            grow, score, n, st["replay_nodes"], st["nodes_evaluated"], st["replay_items"], st["max_noise_ratio"]))
    member = _check_tree_invariants(e, X, grads, depth, grow == "oblivious")
    # predict == bias - lr * value[leaf]
    lrs = np.array([0.1] * d, np.float32) if d == 1 else np.array([0.1] * (d - 1) + [0.05], np.float32)
    exp = -(lrs[None, :] * e["values"][member])
    got = m.predict_numpy(X).reshape(n, d)
    assert np.abs(exp - got).max() <= TOL
    # run-to-run determinism of the integer path at full size
    m2 = configure(make_gpu(**kw), f, d)
    m2.step(X, None, grads)
    e2 = m2.get_ensemble_data()
    for k in ("feature_indices", "feature_values", "values", "edge_weights"):
        assert np.array_equal(e[k], e2[k])
    # the exact tier alone (no replay) must agree except inside the reference's own rounding band
    m3 = configure(make_gpu(tie_replay=False, **kw), f, d)
    m3.step(X, None, grads)
    assert m3.get_ensemble_data()["values"].shape[0] > 0


def test_predict_large_ensemble_is_sum_over_trees():
    n, f, d = 4096, 128, 2
    X, y = synth(n, f, d, 5)
    kw = dict(input_dim=f, output_dim=d, max_depth=6, n_bins=64, split_score_func="cosine", generator_type="quantile",
              batch_size=n, grow_policy="oblivious")
    m = configure(make_gpu(**kw), f, d)
    a = GpuAdaptor(m)
    for it in range(12):
        p = a.predict(X).reshape(n, d)
        a.step(X, (p - y).astype(np.float32))
    full = a.predict(X).astype(np.float64)
    parts = sum(a.predict(X, t, t + 1).astype(np.float64) for t in range(12))   # each = bias(0) - lr*v_t
    assert np.abs(full - parts).max() <= 1e-5


def _numpy_oblivious_predict(e, X, lrs, bias):
    n = X.shape[0]
    D = e["values"].shape[1]
    out = np.tile(np.asarray(bias, np.float64), (n, 1))
    for t in range(e["tree_indices"].shape[0]):
        dep = int(e["depths"][t])
        li = np.zeros(n, np.int64)
        for k in range(dep):
            li |= (X[:, e["feature_indices"][t, k]] > e["feature_values"][t, k]).astype(np.int64) << (dep - 1 - k)
        out -= lrs[None, :] * e["values"][e["tree_indices"][t] + li].astype(np.float64)
    return out


def test_chunked_predict_rollout_shape():
    """BASELINE config 4 shape (scaled): many trees x few observations goes through predict_chunk_kernel; it must
    agree with a float64 evaluation of the same ensemble and with the sequential kernel (large-N path)."""
    rng = np.random.default_rng(3)
    n_trees, depth, f, d = 3000, 6, 128, 2
    nl = n_trees << depth
    e = {"tree_indices": (np.arange(n_trees, dtype=np.int32) << depth), "depths": np.full(n_trees, depth, np.int32),
         "values": (0.01 * rng.standard_normal((nl, d))).astype(np.float32),
         "feature_indices": rng.integers(0, f, (n_trees, depth)).astype(np.int32),
         "feature_values": rng.standard_normal((n_trees, depth)).astype(np.float32),
         "edge_weights": np.zeros((nl, depth), np.float32), "inequality_directions": np.zeros((nl, depth), bool)}
    kw = dict(input_dim=f, output_dim=d, max_depth=depth, n_bins=256, split_score_func="cosine", generator_type="quantile",
              batch_size=8192, grow_policy="oblivious")
    m = configure(make_gpu(**kw), f, d, lrs=[(0.1, 0, 1), (0.01, 1, 2)], bias=[0.25, -0.5])
    m._set_ensemble(e, f)
    X = rng.standard_normal((8192, f)).astype(np.float32)
    got = m.predict_numpy(X)                                  # chunked path
    exp = _numpy_oblivious_predict(e, X, np.array([0.1, 0.01]), [0.25, -0.5])
    assert np.abs(got - exp).max() <= 1e-5
    Xbig = np.concatenate([X] * 5, 0)                          # 40960 rows -> sequential kernel
    seq = m.predict_numpy(Xbig)[:8192]
    assert np.abs(seq.astype(np.float64) - got).max() <= 1e-5
    part = m.predict_numpy(X, 100, 2900)                       # tree sub-range through the chunked path
    e2 = dict(e); e2["tree_indices"] = e["tree_indices"][100:2900]; e2["depths"] = e["depths"][100:2900]
    e2["feature_indices"] = e["feature_indices"][100:2900]; e2["feature_values"] = e["feature_values"][100:2900]
    exp2 = _numpy_oblivious_predict(e2, X, np.array([0.1, 0.01]), [0.25, -0.5])
    assert np.abs(part - exp2).max() <= 1e-5


def test_two_gpu_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29731", os.path.join(root, "tests", "dist_parity_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_PARITY_OK" in r.stdout
