"""J2 -- "drops into GBRL_SB3 unchanged": the reference's UNMODIFIED Python package (gbrl/__init__.py loader, gbrl.learners,
gbrl.models) runs on top of the B200 engine and reproduces what it does on top of its own CPU module.

oracle/Makefile `refpy` installs two copies of the reference's pure-Python package under the git-ignored oracle/_ref/:
site_ref (with the reference's compiled CPU module) and site_b200 (same files, with gbrl_b200/lib/gbrl_cpp*.so -- the
module gbrl/__init__.py:40-118 globs for -- next to them).  tests/learner_worker.py drives GBTLearner,
SharedActorCriticLearner, GBTModel.fit and an autograd-driven ActorCritic on each and dumps ensembles + predictions.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SITE = os.path.join(ROOT, "oracle", "_ref", "site_ref", "gbrl")
B200_SITE = os.path.join(ROOT, "oracle", "_ref", "site_b200", "gbrl")
INT_KEYS = ("tree_indices", "depths", "feature_indices", "inequality_directions")
TOL = 1e-5


def _have_sites():
    return os.path.isdir(REF_SITE) and os.path.isdir(B200_SITE) and any(
        f.startswith("gbrl_cpp") and f.endswith(".so") for f in os.listdir(B200_SITE))


def test_reference_loader_picks_up_the_b200_module():
    """No GPU needed: the reference's own gbrl/__init__.py, unmodified, resolves GBRL_CPP to the B200 engine's class."""
    if not _have_sites():
        pytest.skip("oracle/_ref/site_* not installed (make -C oracle refpy)")
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import gbrl, gbrl_b200\n"
            "assert gbrl.GBRL_CPP is gbrl_b200.GBRL, gbrl.GBRL_CPP\n"
            "from gbrl.learners.gbt_learner import GBRL_CPP as L\n"
            "assert L is gbrl_b200.GBRL\n"
            "print('LOADER_OK', gbrl.__version__)\n") % (ROOT, os.path.dirname(B200_SITE))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "LOADER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_reference_package_files_are_identical_in_both_sites():
    if not _have_sites():
        pytest.skip("oracle/_ref/site_* not installed (make -C oracle refpy)")
    for dp, _, files in os.walk(REF_SITE):
        for fn in files:
            if fn.endswith(".py"):
                a = os.path.join(dp, fn)
                b = os.path.join(B200_SITE, os.path.relpath(a, REF_SITE))
                assert open(a, "rb").read() == open(b, "rb").read(), fn


@pytest.mark.gpu
def test_reference_learners_and_models_run_unchanged_on_the_b200_engine(tmp_path):
    if not _have_sites():
        pytest.skip("oracle/_ref/site_* not installed (make -C oracle refpy)")
    outs = {}
    for eng in ("ref", "b200"):
        o = str(tmp_path / ("%s.npz" % eng))
        env = dict(os.environ, OMP_NUM_THREADS="4")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "learner_worker.py"), eng, o], capture_output=True,
                           text=True, timeout=900, env=env)
        assert r.returncode == 0 and "LEARNER_WORKER_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
        outs[eng] = np.load(o)
    a, b = outs["ref"], outs["b200"]
    assert set(a.files) == set(b.files)
    for k in a.files:
        x, y = a[k], b[k]
        assert x.shape == y.shape, (k, x.shape, y.shape)
        if any(k.endswith(s) for s in INT_KEYS) or k.endswith("ntrees") or k.endswith("meta"):
            assert np.array_equal(x.astype(np.int64), y.astype(np.int64)), k
        elif k.endswith("feature_values"):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)) or np.array_equal(x, y), k   # thresholds: bit-exact
        else:
            err = np.abs(x.astype(np.float64) - y.astype(np.float64)).max() if x.size else 0.0
            assert err <= TOL * max(1.0, float(np.abs(x).max()) if k.endswith("loss") else 1.0), (k, err)
