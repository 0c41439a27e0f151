"""GPU suite (-m gpu): the CUDA engine, called through the C-ABI, against the oracle on the same seeded inputs
and against the golden vectors generated from the reference.

Bar (BASELINE.json north_star): bit-exact split features / thresholds / directions / leaf assignment
(tree_indices, depths), |.| <= 1e-5 on fp32 leaf values and predictions.
"""
import numpy as np
import pytest

from helpers import (GpuAdaptor, OracleAdaptor, boosting_loop, compare_ensembles, configure, default_lrs, make_gpu,
                     make_oracle, run_golden, synth, TOL)

pytestmark = pytest.mark.gpu

GOLDEN = ["greedy_l2", "greedy_cos_ac", "obl_cos_ac", "obl_l2_uniform", "fit_greedy_l2_mb", "fit_obl_cos"]


def _log_stats(tag, st):
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/noise_stats.txt", "a") as fh:
        fh.write("%s | replay_nodes %d / %d evaluated, items %d, max_noise_ratio %.3f\n" % (
            tag, st["replay_nodes"], st["nodes_evaluated"], st["replay_items"], st["max_noise_ratio"]))


def _pair(ref_threads=1, fw=None, lrs=None, **kw):
    f, d = kw["input_dim"], kw["output_dim"]
    o = configure(make_oracle(ref_threads=ref_threads, **kw), f, d, lrs=lrs, fw=fw)
    g = configure(make_gpu(ref_threads=ref_threads, **kw), f, d, lrs=lrs, fw=fw)
    return OracleAdaptor(o), GpuAdaptor(g)


@pytest.mark.parametrize("name", GOLDEN)
def test_gpu_matches_reference_golden(name):
    run_golden(name, make_gpu, GpuAdaptor)


CASES = [
    # n, f, d, depth, bins, score, grow, gen, iters, ref_threads
    (4000, 16, 1, 4, 256, "L2", "oblivious", "quantile", 5, 1),      # BASELINE config 1 shape (reduced N)
    (4000, 16, 1, 4, 256, "cosine", "oblivious", "quantile", 5, 1),
    (6000, 24, 1, 6, 256, "L2", "greedy", "quantile", 4, 4),         # config 2 family: greedy d6 L2 quantile
    (5000, 20, 2, 6, 128, "cosine", "oblivious", "quantile", 3, 1),  # config 3 family: shared policy+value, cosine
    (3000, 40, 3, 5, 64, "cosine", "greedy", "quantile", 4, 1),      # F spans two tiles, D = 3
    (2500, 33, 5, 4, 32, "L2", "greedy", "uniform", 3, 2),           # ragged F, D > 3 (two histogram passes)
    (1800, 7, 2, 5, 200, "L2", "oblivious", "uniform", 4, 3),
    (9000, 12, 1, 7, 256, "cosine", "greedy", "quantile", 3, 1),     # deep greedy tree, small leaves
]


@pytest.mark.parametrize("n,f,d,depth,bins,score,grow,gen,iters,T", CASES)
def test_step_parity_vs_oracle(n, f, d, depth, bins, score, grow, gen, iters, T):
    X, y = synth(n, f, d, seed=n + f)
    kw = dict(input_dim=f, output_dim=d, max_depth=depth, n_bins=bins, par_th=10, split_score_func=score,
              generator_type=gen, batch_size=n, grow_policy=grow)
    fw = (1.0 + 0.05 * np.arange(f)).astype(np.float32)
    o, g = _pair(ref_threads=T, fw=fw, **kw)
    boosting_loop([o, g], X, y, iters)
    st = g.m.get_stats()
    assert st["replay_overflow"] == 0
    assert st["n_trees"] == iters
    _log_stats("step n=%d f=%d d=%d depth=%d %s %s T=%d" % (n, f, d, depth, score, grow, T), st)


MEDIUM = [
    # large enough for long float chains (many binade changes, near-ties inside the replay band), small enough for the oracle
    (60000, 12, 1, 5, 256, "L2", "greedy", "quantile", 3, 4),
    (40000, 10, 2, 4, 256, "cosine", "oblivious", "quantile", 3, 4),
    (50000, 8, 1, 6, 256, "cosine", "greedy", "quantile", 2, 1),
    (30000, 6, 3, 4, 128, "L2", "greedy", "quantile", 2, 2),          # output_dim 3: one CTA per replay item
]


@pytest.mark.parametrize("n,f,d,depth,bins,score,grow,gen,iters,T", MEDIUM)
def test_step_parity_medium(n, f, d, depth, bins, score, grow, gen, iters, T):
    X, y = synth(n, f, d, seed=n + f)
    kw = dict(input_dim=f, output_dim=d, max_depth=depth, n_bins=bins, par_th=10, split_score_func=score,
              generator_type=gen, batch_size=n, grow_policy=grow)
    o, g = _pair(ref_threads=T, **kw)
    boosting_loop([o, g], X, y, iters)
    st = g.m.get_stats()
    assert st["replay_overflow"] == 0
    _log_stats("medium n=%d f=%d d=%d depth=%d %s %s T=%d chain fast/slow %d/%d" % (
        n, f, d, depth, score, grow, T, st["chain_blocks_fast"], st["chain_blocks_slow"]), st)


@pytest.mark.parametrize("case", [CASES[2], CASES[3], MEDIUM[0], MEDIUM[1]])
def test_step_parity_kernel_variants(case):
    """The comparison variants stay correct: per-item histogram kernel (hist_variant=1) and one CTA per replay item
    (replay_variant=1) against the oracle, like the default kernels."""
    n, f, d, depth, bins, score, grow, gen, iters, T = case
    X, y = synth(n, f, d, seed=n + f + 2)
    kw = dict(input_dim=f, output_dim=d, max_depth=depth, n_bins=bins, par_th=10, split_score_func=score,
              generator_type=gen, batch_size=n, grow_policy=grow)
    o = configure(make_oracle(ref_threads=T, **kw), f, d)
    g = configure(make_gpu(ref_threads=T, hist_variant=1, replay_variant=1, **kw), f, d)
    boosting_loop([OracleAdaptor(o), GpuAdaptor(g)], X, y, min(iters, 2))


def test_step_parity_direct_replay_items():
    """Replay items whose side-bit plane does not fit the stream buffer are gathered by their own CTA
    (replay_par_kernel); GBRL_B200_REPLAY_DIRECT=1 sends every item down that path.  Runs in a subprocess because the
    switch is latched when the library first replays."""
    import subprocess, sys, os
    code = (
        "import os, sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from helpers import *\n"
        "import test_gpu_parity as t\n"
        "for case in (t.MEDIUM[0], t.MEDIUM[2], t.CASES[7]):\n"
        "    n, f, d, depth, bins, score, grow, gen, iters, T = case\n"
        "    X, y = synth(n, f, d, seed=n + f)\n"
        "    kw = dict(input_dim=f, output_dim=d, max_depth=depth, n_bins=bins, par_th=10, split_score_func=score,\n"
        "              generator_type=gen, batch_size=n, grow_policy=grow)\n"
        "    o, g = t._pair(ref_threads=T, **kw)\n"
        "    boosting_loop([o, g], X, y, 2)\n"
        "    assert g.m.get_stats()['replay_items'] > 0\n"
        "print('DIRECT_REPLAY_OK')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GBRL_B200_REPLAY_DIRECT="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "DIRECT_REPLAY_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


@pytest.mark.parametrize("case", [CASES[2], CASES[3], CASES[7]])
def test_step_parity_feature_major_codes(case, monkeypatch):
    """Large matrices take the side of a split from the feature-major u16 code copy (x > thr[f][j] <=> code > j) instead
    of the fp32 value; force that path on small inputs and check it against the oracle like any other."""
    monkeypatch.setenv("GBRL_B200_FEATMAJOR", "1")
    n, f, d, depth, bins, score, grow, gen, iters, T = case
    X, y = synth(n, f, d, seed=n + f + 1)
    kw = dict(input_dim=f, output_dim=d, max_depth=depth, n_bins=bins, par_th=10, split_score_func=score,
              generator_type=gen, batch_size=n, grow_policy=grow)
    o, g = _pair(ref_threads=T, **kw)
    boosting_loop([o, g], X, y, iters)
    assert g.m.get_stats()["replay_overflow"] == 0


@pytest.mark.parametrize("grow,score", [("greedy", "L2"), ("oblivious", "cosine")])
def test_fit_parity_minibatch_feature_major(grow, score, monkeypatch):
    monkeypatch.setenv("GBRL_B200_FEATMAJOR", "1")
    test_fit_parity_minibatch(grow, score)


@pytest.mark.parametrize("grow,score", [("greedy", "L2"), ("oblivious", "cosine")])
def test_fit_parity_minibatch(grow, score):
    n, f, d = 3100, 10, 2
    X, y = synth(n, f, d, 11)
    kw = dict(input_dim=f, output_dim=d, max_depth=4, n_bins=64, par_th=10, split_score_func=score,
              generator_type="quantile", batch_size=1000, grow_policy=grow)
    o, g = _pair(ref_threads=2, **kw)
    lo, lg = o.fit(X, y, 7), g.fit(X, y, 7)
    compare_ensembles(o.ensemble(), g.ensemble(), "fit")
    assert abs(lo - lg) <= 1e-5 * max(1.0, abs(lo))
    assert np.abs(o.predict(X).astype(np.float64) - g.predict(X)).max() <= TOL
    assert np.abs(o.o.get_ensemble_data()["values"]).max() > 0


def test_candidates_match_oracle_bitwise():
    n, f = 5000, 11
    X, y = synth(n, f, 1, 2)
    X[:, 3] = np.round(X[:, 3])          # heavy ties -> duplicated quantile thresholds
    X[:, 5] = 1.5                        # constant column
    kw = dict(input_dim=f, output_dim=1, max_depth=2, n_bins=256, split_score_func="L2", generator_type="quantile",
              batch_size=n, grow_policy="greedy")
    o, g = _pair(**kw)
    grads = (0.0 - y).astype(np.float32)
    _, thr = o.o.root_scores(X, grads)
    g.step(X, grads)
    got = g.m.get_candidates()
    # bit for bit, except the sign of a zero threshold: -0.0 and +0.0 compare equal, so which of the two a sort leaves at a given
    # rank is an artefact of the sort (std::sort in the reference, qsort in the oracle, a radix select here) and no comparison
    # x > threshold can tell them apart
    nz = (thr != 0) | (got != 0)
    assert np.array_equal(thr.view(np.uint32)[nz], got.view(np.uint32)[nz])
    assert np.array_equal(thr, got)
    o.step(X, grads)
    compare_ensembles(o.ensemble(), g.ensemble(), "ties")


def test_exact_tier_scores_close_to_reference_scores():
    """The exact-arithmetic scores must sit within the rounding band of the reference's sequential-fp32 scores
    (this is what the near-tie replay band is calibrated on)."""
    n, f = 20000, 8
    X, y = synth(n, f, 2, 4)
    kw = dict(input_dim=f, output_dim=2, max_depth=1, n_bins=64, split_score_func="cosine", generator_type="quantile",
              batch_size=n, grow_policy="greedy")
    o, g = _pair(**kw)
    grads = (0.3 - y).astype(np.float32)
    sref, _ = o.o.root_scores(X, grads)
    g.step(X, grads)
    sgpu = g.m.get_root_scores()
    rel = np.abs(sref.astype(np.float64) - sgpu) / np.abs(sref).max()
    assert rel.max() < 4 * 2.0 ** -24 * np.sqrt(n)


def test_subtraction_trick_and_replay_do_not_change_trees():
    n, f, d = 12000, 20, 1
    X, y = synth(n, f, d, 9)
    base = dict(input_dim=f, output_dim=d, max_depth=6, n_bins=256, split_score_func="L2", generator_type="quantile",
                batch_size=n, grow_policy="greedy")
    a = GpuAdaptor(configure(make_gpu(use_subtraction=True, **base), f, d))
    b = GpuAdaptor(configure(make_gpu(use_subtraction=False, **base), f, d))
    boosting_loop([a, b], X, y, 3)       # integer histograms: parent - sibling == direct, bit for bit
    ea, eb = a.ensemble(), b.ensemble()
    assert np.array_equal(ea["values"], eb["values"])


def test_determinism_run_to_run():
    n, f, d = 8000, 18, 2
    X, y = synth(n, f, d, 21)
    kw = dict(input_dim=f, output_dim=d, max_depth=5, n_bins=128, split_score_func="cosine", generator_type="quantile",
              batch_size=n, grow_policy="oblivious")
    outs = []
    for _ in range(2):
        m = GpuAdaptor(configure(make_gpu(**kw), f, d))
        for it in range(3):
            p = m.predict(X).reshape(n, d)
            m.step(X, (p - y).astype(np.float32))
        outs.append((m.ensemble(), m.predict(X)))
    for k in ("values", "feature_values", "feature_indices", "edge_weights"):
        assert np.array_equal(outs[0][0][k], outs[1][0][k])
    assert np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("grow,score", [("greedy", "L2"), ("oblivious", "cosine"), ("greedy", "cosine"), ("oblivious", "L2")])
def test_discrete_features_and_exact_ties(grow, score):
    """Small-integer features, a constant column and an exact copy of a column: most quantile thresholds are duplicates and many
    candidates describe the SAME partition, so the arg-max is decided by the reference's rule for exact ties (lowest candidate
    index, fitter.cpp:338-356 / 427-445) and by the path guard (node.cpp:154-166), not by score differences."""
    n, f, d = 6000, 9, 2
    rng = np.random.default_rng(5)
    X = rng.integers(0, 4, size=(n, f)).astype(np.float32)
    X[:, 2] = 7.0                         # constant column: n_bins equal thresholds, every candidate has an empty side
    X[:, 6] = X[:, 1]                     # exact copy: candidates of column 6 tie with those of column 1
    X[:, 4] = rng.standard_normal(n).astype(np.float32)
    y = (np.stack([X[:, 1] - X[:, 3], X[:, 4] * (X[:, 0] > 1)], 1) + 0.05 * rng.standard_normal((n, d))).astype(np.float32)
    kw = dict(input_dim=f, output_dim=d, max_depth=5, n_bins=64, par_th=10, split_score_func=score, generator_type="quantile",
              batch_size=n, grow_policy=grow)
    o, g = _pair(**kw)
    boosting_loop([o, g], X, y, 4)
    assert g.m.get_stats()["replay_overflow"] == 0


def test_targets_with_huge_dynamic_range():
    """Gradients spanning 12 decades and exactly-zero gradients: the fixed-point scale of the histogram is taken from max |g|, so
    tiny gradients quantise to 0 in the exact tier -- the replay tier and the leaf values (raw gradients) must still be the reference's."""
    n, f, d = 5000, 8, 1
    X, y = synth(n, f, d, 77)
    y = y.copy()
    y[::7] *= 1e6
    y[1::7] *= 1e-6
    y[2::7] = 0.0
    kw = dict(input_dim=f, output_dim=d, max_depth=4, n_bins=128, par_th=10, split_score_func="L2", generator_type="quantile",
              batch_size=n, grow_policy="greedy")
    o, g = _pair(**kw)
    n_, d_ = y.shape
    for it in range(3):                    # leaf values reach 1e5: compare relative to their magnitude
        po = np.asarray(o.predict(X)).reshape(n_, d_)
        grad = (po - y).astype(np.float32)
        o.step(X, grad); g.step(X, grad)
        eo, eg = o.ensemble(), g.ensemble()
        for k in ("tree_indices", "depths", "feature_indices", "inequality_directions"):
            assert np.array_equal(np.asarray(eo[k]).astype(np.int64), np.asarray(eg[k]).astype(np.int64)), (it, k)
        assert np.array_equal(np.asarray(eo["feature_values"]), np.asarray(eg["feature_values"]))
        vo, vg = np.asarray(eo["values"], np.float64), np.asarray(eg["values"], np.float64).reshape(np.asarray(eo["values"]).shape)
        assert np.all(np.abs(vo - vg) <= TOL * np.maximum(1.0, np.abs(vo))), np.abs(vo - vg).max()


@pytest.mark.parametrize("d,score,grow", [(20, "cosine", "greedy"), (48, "L2", "greedy"), (48, "cosine", "oblivious"), (64, "L2", "oblivious")])
def test_wide_outputs(d, score, grow):
    """output_dim beyond 16: the split scan is instantiated for 32 and 64 output dimensions (split.cu launch_scan), the histogram takes
    two dimensions per pass, the replay runs one lane per output dimension."""
    n, f = 2500, 7
    X, y = synth(n, f, d, seed=d)
    kw = dict(input_dim=f, output_dim=d, max_depth=3, n_bins=48, par_th=10, split_score_func=score, generator_type="quantile",
              batch_size=n, grow_policy=grow)
    o, g = _pair(ref_threads=2, lrs=[(0.1, 0, d)], **kw)
    boosting_loop([o, g], X, y, 2)
    assert g.m.get_stats()["replay_overflow"] == 0


@pytest.mark.parametrize("n", [1, 2, 37, 300])
def test_tiny_inputs(n):
    """n_samples below n_bins+1 (quantile ranks collapse), single-sample nodes, empty children."""
    f, d = 5, 1
    X, y = synth(max(n, 2), f, d, 31)
    X, y = X[:n], y[:n]
    kw = dict(input_dim=f, output_dim=d, max_depth=3, n_bins=256, split_score_func="L2", generator_type="uniform",
              batch_size=max(n, 1), grow_policy="greedy")
    o, g = _pair(**kw)
    boosting_loop([o, g], X, y, 2)


def test_min_data_in_leaf_and_feature_weights():
    n, f, d = 3000, 6, 1
    X, y = synth(n, f, d, 41)
    kw = dict(input_dim=f, output_dim=d, max_depth=5, n_bins=64, split_score_func="L2", generator_type="quantile",
              batch_size=n, grow_policy="greedy", min_data_in_leaf=200)
    fw = np.array([1, 0.5, 2, 1, 0.25, 1], np.float32)
    o, g = _pair(fw=fw, **kw)
    boosting_loop([o, g], X, y, 3)


def test_predict_tree_ranges_and_loaded_ensemble():
    n, f, d = 2000, 9, 2
    X, y = synth(n, f, d, 51)
    for grow in ("greedy", "oblivious"):
        kw = dict(input_dim=f, output_dim=d, max_depth=4, n_bins=32, split_score_func="cosine", generator_type="quantile",
                  batch_size=n, grow_policy=grow)
        o, g = _pair(**kw)
        boosting_loop([o, g], X, y, 4, check_each=False)
        for (a, b) in ((0, 0), (1, 3), (2, 4), (0, 1)):
            assert np.abs(o.predict(X, a, b).astype(np.float64) - g.predict(X, a, b)).max() <= TOL
        # an ensemble in the reference layout loaded through the C-ABI predicts identically
        g2 = configure(make_gpu(**kw), f, d)
        g2._set_ensemble(o.ensemble(), f)
        assert np.abs(o.predict(X).astype(np.float64) - g2.predict_numpy(X)).max() <= TOL
        with pytest.raises(RuntimeError):
            g.m.predict_tensor(X, 0, 99)


def test_device_inputs_and_dlpack_output():
    import torch
    n, f, d = 1500, 8, 2
    X, y = synth(n, f, d, 61)
    kw = dict(input_dim=f, output_dim=d, max_depth=3, n_bins=32, split_score_func="cosine", generator_type="quantile",
              batch_size=n, grow_policy="greedy")
    o, g = _pair(**kw)
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    for it in range(2):
        p = torch.from_dlpack(g.m.predict((Xd.data_ptr(), tuple(Xd.shape), "torch.float32", "cuda"), None))
        assert p.is_cuda and tuple(p.shape) == (n, d)
        gr = (p - yd).contiguous()
        g.m.step((Xd.data_ptr(), tuple(Xd.shape), "torch.float32", "cuda"), None,
                 (gr.data_ptr(), tuple(gr.shape), "torch.float32", "cuda"))
        o.step(X, gr.cpu().numpy())
    compare_ensembles(o.ensemble(), g.ensemble(), "device inputs")


def test_error_behaviour_matches_reference():
    n, f, d = 100, 4, 2
    X, y = synth(n, f, d, 71)
    m = make_gpu(input_dim=f, output_dim=d, max_depth=2, n_bins=8, split_score_func="L2", grow_policy="greedy", batch_size=n)
    with pytest.raises(RuntimeError):            # gbrl.cpp:463 start >= stop
        m.set_optimizer("SGD", "const", 0.1, 1, 1)
    with pytest.raises(RuntimeError):            # gbrl.cpp:468 out of range
        m.set_optimizer("SGD", "const", 0.1, 0, 3)
    configure(m, f, d)
    with pytest.raises(RuntimeError):            # gbrl.cpp:457 limit = output_dim optimizers
        m.set_optimizer("SGD", "const", 0.1, 0, 1)
    with pytest.raises(RuntimeError):            # binding.cpp:473 gradient dim
        m.step(X, None, y[:, :1].copy())
    with pytest.raises(RuntimeError):            # binding.cpp:517 feature count
        m.step(X[:, :3].copy(), None, y)
    with pytest.raises(RuntimeError):            # binding.cpp:493 sample count
        m.step(X[:50].copy(), None, y)
    assert m.get_num_trees() == 0
    p = m.predict_numpy(X)                       # no trees: bias only (predictor.cpp:125-128)
    assert np.array_equal(p, np.zeros((n, d), np.float32))


def test_checkpoint_roundtrip_with_reference(tmp_path):
    """a model trained on the GPU saves in the reference's `.gbrl_model` format; the reference loads it on
    CPU and predicts the same; the file the reference writes loads back into the GPU engine."""
    from oracle.oracle import load_reference
    ref = load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    from gbrl_b200 import GBRL
    n, f, d = 2000, 9, 2
    X, y = synth(n, f, d, 81)
    for grow in ("greedy", "oblivious"):
        kw = dict(input_dim=f, output_dim=d, max_depth=4, n_bins=64, split_score_func="cosine", generator_type="quantile",
                  batch_size=n, grow_policy=grow)
        g = GpuAdaptor(configure(make_gpu(**kw), f, d, bias=[0.2, -0.3]))
        for it in range(3):
            p = g.predict(X).reshape(n, d)
            g.step(X, (p - y).astype(np.float32))
        path = str(tmp_path / ("gpu_%s.gbrl_model" % grow))
        assert g.m.save(path) == 0
        r = ref.GBRL.load(path)
        pr = np.array(r.predict(X, None), copy=True).reshape(n, d)
        assert np.abs(pr.astype(np.float64) - g.predict(X).reshape(n, d)).max() <= TOL
        path2 = str(tmp_path / ("ref_%s.gbrl_model" % grow))
        assert r.save(path2) == 0
        g2 = GBRL.load(path2)
        assert g2.get_num_trees() == 3
        assert np.abs(g2.predict_numpy(X).reshape(n, d).astype(np.float64) - pr).max() <= TOL
        # training continues after a load (same trees as continuing the original model)
        p = g.predict(X).reshape(n, d)
        grads = (p - y).astype(np.float32)
        g.step(X, grads)
        g2.step(X, None, grads)
        compare_ensembles(g.ensemble(), g2.get_ensemble_data(), "continue after load")
        test_checkpoint_roundtrip_with_reference.keep = getattr(test_checkpoint_roundtrip_with_reference, "keep", []) + [r]
