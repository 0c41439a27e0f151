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
# fit(shuffle=True) with mini-batches: the ranks must draw the SAME permutation (shared seed, capi.cu) -> identical ensembles
import hashlib  # noqa: E402
n, f, d = 40000, 40, 1
X, y = synth(n, f, d, 11)
m = configure(GBRL(input_dim=f, output_dim=d, policy_dim=d, max_depth=4, n_bins=64, split_score_func="cosine", generator_type="quantile",
                   batch_size=10000, grow_policy="oblivious", ref_threads=1, device="cuda:%d" % local), f, d)
m.init_distributed()
m.fit(X, None, y, 2, True, "MultiRMSE")
e = m.get_ensemble_data()
h = hashlib.sha256(b"".join(np.ascontiguousarray(e[k]).tobytes() for k in ("feature_indices", "feature_values", "values"))).digest()[:8]
mine = torch.tensor([int.from_bytes(h, "little", signed=True)], device="cuda", dtype=torch.int64)
allh = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allh, mine)
same = all(int(t.item()) == int(allh[0].item()) for t in allh)
if not same:
    print("rank %d: shuffled fit differs across ranks" % rank)
ok = ok and same and e["values"].shape[0] > 0
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t)
if rank == 0 and int(t.item()) == world:
    print("DIST_PARITY_OK")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
