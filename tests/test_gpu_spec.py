"""Speculative levels: the replay chains run off the critical path, a level whose decision they change is returned to
(gbrl_b200/csrc/tree.cu grow_tree).  The result must be the reference's either way."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, os.path.join(HERE, "spec_worker.py")], capture_output=True, text=True, timeout=1200, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SPEC_OK" in r.stdout
    return r.stdout


def test_speculative_levels_match_oracle_and_synchronous_replay():
    _run({})


def test_forced_wrong_speculation_is_rolled_back():
    out = _run({"GBRL_B200_SPEC_FORCE_FLIP": "1"})
    assert "rollbacks=0" not in out
