"""torchrun worker of tests/test_gpu_scale.py::test_two_gpu_matches_single_gpu: the feature-sharded, all-reduced
histogram must give the same trees as a single GPU (integer sums: bit-identical)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import configure, synth  # noqa: E402
from gbrl_b200 import GBRL  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
ok = True
# F=96 -> 3 feature tiles over 2 ranks (tile sharding); F=20 -> 1 tile, so the 2 ranks split the row chunks instead
for (n, f, d, depth, score, grow) in ((60000, 96, 2, 5, "cosine", "greedy"), (50000, 20, 1, 4, "L2", "oblivious")):
    X, y = synth(n, f, d, 7)
    kw = dict(input_dim=f, output_dim=d, policy_dim=d, max_depth=depth, n_bins=128, split_score_func=score,
              generator_type="quantile", batch_size=n, grow_policy=grow, ref_threads=1, device="cuda:%d" % local)
    single = configure(GBRL(**kw), f, d)
    sharded = configure(GBRL(**kw), f, d)
    sharded.init_distributed()
    for it in range(3):
        p = single.predict_numpy(X).reshape(n, d)
        g = (p - y).astype(np.float32)
        single.step(X, None, g)
        sharded.step(X, None, g)
    a, b = single.get_ensemble_data(), sharded.get_ensemble_data()
    ok = ok and all(np.array_equal(a[k], b[k]) for k in ("tree_indices", "depths", "feature_indices", "feature_values", "values", "edge_weights"))
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t)
if rank == 0 and int(t.item()) == world:
    print("DIST_PARITY_OK")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
