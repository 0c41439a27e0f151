"""CPU suite (-m "not gpu"): pins the oracle.

 * oracle/gbrl_oracle.c (plain-C restatement) against the golden vectors generated from the reference;
 * the restatement against the compiled reference itself (oracle/_ref) when it is present, bit for bit,
   at the OpenMP thread count of this process (conftest pins OMP_NUM_THREADS);
 * reference quirks the engine has to mirror.
"""
import numpy as np
import pytest

from conftest import REF_THREADS
from helpers import (OracleAdaptor, compare_ensembles, configure, default_lrs, make_oracle, run_golden, synth)

GOLDEN = ["greedy_l2", "greedy_cos_ac", "obl_cos_ac", "obl_l2_uniform", "fit_greedy_l2_mb", "fit_obl_cos"]


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_matches_golden(name):
    run_golden(name, make_oracle, OracleAdaptor)


def _reference():
    from oracle.oracle import load_reference
    return load_reference()


@pytest.mark.parametrize("score", ["L2", "cosine"])
@pytest.mark.parametrize("grow", ["greedy", "oblivious"])
def test_oracle_matches_compiled_reference(score, grow):
    ref = _reference()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    from oracle.oracle import make_reference
    n, f, d, depth = 2500, 9, 2, 5
    X, y = synth(n, f, d, 3)
    kw = dict(input_dim=f, output_dim=d, max_depth=depth, n_bins=128, par_th=10, split_score_func=score,
              generator_type="quantile", batch_size=n, grow_policy=grow)
    lrs = default_lrs(d)
    r = make_reference(ref, lrs=lrs, **kw)
    o = configure(make_oracle(ref_threads=REF_THREADS, **kw), f, d, lrs=lrs)
    for it in range(4):
        pr = np.array(r.predict(X, None), copy=True).reshape(n, d)
        po = np.asarray(o.predict(X)).reshape(n, d)
        assert np.array_equal(pr, po), "iteration %d predictions differ" % it
        g = (pr - y).astype(np.float32)
        r.step(X, None, g)
        o.step(X, g)
    er = r.get_ensemble_data()          # once per model (reference capsule quirk)
    compare_ensembles(er, o.get_ensemble_data(), "%s/%s" % (score, grow), tol=0.0)
    test_oracle_matches_compiled_reference.keep = getattr(test_oracle_matches_compiled_reference, "keep", []) + [er, r]


def test_quantile_duplicates_are_kept():
    """split_candidate_generator.cpp:241: the dedup branch is dead on CPU -> exactly n_bins thresholds per
    feature, duplicates included (a constant column yields n_bins equal thresholds)."""
    from oracle.oracle import Oracle
    X = np.zeros((300, 2), np.float32)
    X[:, 1] = np.arange(300)
    o = Oracle(input_dim=2, output_dim=1, max_depth=2, n_bins=16, split_score_func="L2")
    scores, thr = o.root_scores(X, X[:, 1:2].copy())
    assert thr.shape == (32,)
    assert np.all(thr[:16] == 0.0)
    assert np.all(np.diff(thr[16:]) > 0)


def test_multirmse_tail_quirk():
    """loss.cpp:34-62: with T threads only T*(n_elements//T) gradients are written."""
    n, f = 443, 3   # 443 elements, T = min(4, 44) -> 4 threads * 110 = 440 written
    X, y = synth(n, f, 1, 5)
    a = configure(make_oracle(ref_threads=4, input_dim=f, output_dim=1, max_depth=2, n_bins=8, batch_size=n, split_score_func="L2",
                              generator_type="uniform", grow_policy="greedy"), f, 1)
    b = configure(make_oracle(ref_threads=1, input_dim=f, output_dim=1, max_depth=2, n_bins=8, batch_size=n, split_score_func="L2",
                              generator_type="uniform", grow_policy="greedy"), f, 1)
    la, lb = a.fit(X, y, 3), b.fit(X, y, 3)
    assert np.isfinite(la) and np.isfinite(lb)
    # the thread-partitioned run must differ from the serial one somewhere (tail gradients stay 0)
    va, vb = a.get_ensemble_data()["values"], b.get_ensemble_data()["values"]
    assert va.shape != vb.shape or not np.array_equal(va, vb)
