"""Worker of tests/test_gpu_spec.py (own process: GBRL_B200_SPEC_FORCE_FLIP is read once per process).

Speculative levels (tree.cu grow_tree) must never change a result: for every case the engine with speculative levels is
compared, over a boosting loop on the same gradient stream, with the oracle AND with the same engine waiting for every level's
replay (replay_variant bit 1).  With GBRL_B200_SPEC_FORCE_FLIP=1 the speculative decision deliberately takes a wrong candidate
wherever a node has replay items, so every such level is verified, returned to and decided again."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import GpuAdaptor, OracleAdaptor, boosting_loop, compare_ensembles, configure, make_gpu, make_oracle, synth  # noqa: E402

CASES = [
    # n, f, d, depth, bins, score, grow, iters, ref_threads, replay_variant (bit 0)
    (30000, 24, 1, 6, 256, "L2", "greedy", 4, 4, 0),
    (30000, 20, 2, 6, 128, "cosine", "oblivious", 4, 1, 0),
    (20000, 16, 1, 5, 256, "cosine", "oblivious", 4, 2, 0),
    (12000, 40, 3, 5, 64, "cosine", "greedy", 3, 1, 0),       # D = 3: per-item chain kernels
    (8000, 33, 6, 4, 32, "L2", "greedy", 3, 2, 0),            # D > 4: one lane per output dimension
    (20000, 24, 2, 5, 256, "L2", "greedy", 3, 1, 1),          # replay_variant 1: one CTA per replay item
]
forced = os.environ.get("GBRL_B200_SPEC_FORCE_FLIP", "") == "1"
tot_rollbacks = tot_spec = tot_nodes = 0
for (n, f, d, depth, bins, score, grow, iters, T, rv) in CASES:
    X, y = synth(n, f, d, seed=n + f)
    kw = dict(input_dim=f, output_dim=d, max_depth=depth, n_bins=bins, par_th=10, split_score_func=score,
              generator_type="quantile", batch_size=n, grow_policy=grow)
    fw = (1.0 + 0.05 * np.arange(f)).astype(np.float32)
    o = OracleAdaptor(configure(make_oracle(ref_threads=T, **kw), f, d, fw=fw))
    spec = GpuAdaptor(configure(make_gpu(ref_threads=T, replay_variant=rv, **kw), f, d, fw=fw))
    sync = GpuAdaptor(configure(make_gpu(ref_threads=T, replay_variant=rv | 2, **kw), f, d, fw=fw))
    boosting_loop([o, spec, sync], X, y, iters)
    a, b = spec.ensemble(), sync.ensemble()
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), "speculative and synchronous replay differ in %s" % k
    st, st2 = spec.m.get_stats(), sync.m.get_stats()
    assert st["spec_trees"] == iters and st2["spec_trees"] == 0, (st, st2)
    assert st["replay_overflow"] == 0
    tot_rollbacks += st["spec_rollbacks"]; tot_spec += st["spec_trees"]; tot_nodes += st["replay_nodes"]
    print("case", (n, f, d, depth, score, grow, rv), "spec_trees", st["spec_trees"], "rolled back levels", st["spec_rollbacks"],
          "replay nodes", st["replay_nodes"], "flips", st["replay_flips"], flush=True)
assert tot_nodes > 0, "no case had a near-tie: the test does not exercise the replay"
if forced:
    assert tot_rollbacks > 0, "forced wrong speculation was never rolled back"
print("SPEC_OK rollbacks=%d trees=%d" % (tot_rollbacks, tot_spec))
