"""Worker of tests/test_learner_dropin.py: drives the reference's UNMODIFIED Python package (gbrl.learners, gbrl.models)
on one engine and dumps what it produced.

    python tests/learner_worker.py ref  out.npz     # oracle/_ref/site_ref : the reference's own CPU module
    python tests/learner_worker.py b200 out.npz     # oracle/_ref/site_b200: same package files, gbrl_cpp*.so = the B200 shim

The two site directories hold byte-identical copies of gbrl/{__init__,common,learners,models}; only the compiled module
the reference's loader finds next to them differs (oracle/Makefile `refpy`).  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

ENGINE, OUT = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SITE = os.path.join(ROOT, "oracle", "_ref", "site_ref" if ENGINE == "ref" else "site_b200")
sys.path.insert(0, ROOT)
sys.path.insert(0, SITE)
os.environ.setdefault("OMP_NUM_THREADS", "4")

import numpy as np  # noqa: E402
import torch as th  # noqa: E402
import gbrl  # noqa: E402
from gbrl.learners.actor_critic_learner import SharedActorCriticLearner  # noqa: E402
from gbrl.learners.gbt_learner import GBTLearner  # noqa: E402
from gbrl.models.actor_critic import ActorCritic  # noqa: E402
from gbrl.models.gbt import GBTModel  # noqa: E402

assert os.path.dirname(gbrl.__file__).startswith(SITE), gbrl.__file__
backend = getattr(sys.modules.get("gbrl_cpp", gbrl._gbrl_cpp_module), "__backend__", "reference")
assert (backend == "gbrl_b200") == (ENGINE == "b200"), backend
DEV = "cpu" if ENGINE == "ref" else "cuda"
out = {}
KEYS = ("tree_indices", "depths", "feature_indices", "feature_values", "inequality_directions", "edge_weights", "values")


def data(n, f, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, f)).astype(np.float32)
    W = rng.standard_normal((f, d)).astype(np.float32)
    y = (np.tanh(X @ W / np.sqrt(f)) + 0.1 * rng.standard_normal((n, d))).astype(np.float32)
    return X, y


def T(a):
    return th.from_numpy(np.ascontiguousarray(a)).to(DEV)


def dump(tag, cpp_model):
    e = cpp_model.get_ensemble_data()          # the reference allows ONE call per model (its capsules double-free)
    for k in KEYS:
        out["%s_%s" % (tag, k)] = np.array(e[k], copy=True)
    dump.keep.append(e)


dump.keep = []

# ---- S1: GBTLearner.step with torch tensors (gbt_learner.py:105-148), predict (:455-500), save / load (:183-275)
n, f, d = 3000, 12, 1
X, y = data(n, f, d, 7)
ts = {"max_depth": 4, "min_data_in_leaf": 0, "n_bins": 64, "par_th": 10, "batch_size": n, "grow_policy": "greedy"}
opt = {"algo": "SGD", "init_lr": 0.1, "scheduler": "Const", "start_idx": 0, "stop_idx": 1}
L = GBTLearner(input_dim=f, output_dim=d, tree_struct=ts, optimizers=dict(opt), params={"split_score_func": "L2", "generator_type": "Quantile"},
               verbose=0, device=DEV)
L.reset()
L.set_bias(np.zeros(d, np.float32))
Xt, yt = T(X), T(y)
for it in range(3):
    p = L.predict(Xt, requires_grad=False, tensor=True)
    g = (p.reshape(n, d) - yt).contiguous()
    L.step(Xt, g)
out["s1_pred"] = L.predict(Xt, requires_grad=False, tensor=True).detach().cpu().numpy().reshape(n, d)
out["s1_pred_range"] = L.predict(Xt, requires_grad=False, start_idx=1, stop_idx=3, tensor=True).detach().cpu().numpy().reshape(n, d)
out["s1_ntrees"] = np.array([L.get_num_trees(), L.get_iteration()])
path = OUT + ".s1"
L.save(path)
L2 = GBTLearner.load(path, DEV)
out["s1_loaded_pred"] = L2.predict(Xt, requires_grad=False, tensor=True).detach().cpu().numpy().reshape(n, d)
out["s1_loaded_meta"] = np.array([L2.get_num_trees(), L2.get_iteration(), L2.input_dim, L2.output_dim])
g = (L2.predict(Xt, requires_grad=False, tensor=True).reshape(n, d) - yt).contiguous()
L2.step(Xt, g)                                   # training continues after load
out["s1_cont_pred"] = L2.predict(Xt, requires_grad=False, tensor=True).detach().cpu().numpy().reshape(n, d)
dump("s1", L2._cpp_model)

# ---- S2: SharedActorCriticLearner (actor_critic_learner.py:88): one shared tree, policy columns [0, D-1) + value column
n, f, d = 2500, 9, 4
X, y = data(n, f, d, 11)
ts = {"max_depth": 4, "min_data_in_leaf": 0, "n_bins": 48, "par_th": 10, "batch_size": n, "grow_policy": "oblivious"}
po = {"algo": "SGD", "init_lr": 0.1, "scheduler": "Const", "start_idx": 0, "stop_idx": d - 1}
vo = {"algo": "SGD", "init_lr": 0.02, "scheduler": "Const", "start_idx": d - 1, "stop_idx": d}
A = SharedActorCriticLearner(input_dim=f, output_dim=d, tree_struct=ts, policy_optimizer=po, value_optimizer=vo,
                             params={"split_score_func": "Cosine", "generator_type": "Quantile"}, verbose=0, device=DEV)
A.reset()
A.set_bias(np.zeros(d, np.float32))
Xt, yt = T(X), T(y)
for it in range(3):
    pol, val = A.predict(Xt, requires_grad=False, tensor=True)
    pg = (pol - yt[:, :d - 1]).contiguous()
    vg = (val - yt[:, d - 1]).contiguous()
    A.step(Xt, (pg, vg))
pol, val = A.predict(Xt, requires_grad=False, tensor=True)
out["s2_policy"] = pol.detach().cpu().numpy()
out["s2_value"] = val.detach().cpu().numpy()
A2 = A.copy()                                    # GBRL_CPP(model) copy-constructor, binding.cpp:441
pol2, val2 = A2.predict(Xt, requires_grad=False, tensor=True)
out["s2_copy_policy"] = pol2.detach().cpu().numpy()
dump("s2", A._cpp_model)

# ---- S3: GBTModel.fit with NumPy inputs (supervised MultiRMSE loop, fitter.cpp:117-261) and __call__
n, f, d = 2000, 8, 2
X, y = data(n, f, d, 13)
ts = {"max_depth": 3, "min_data_in_leaf": 0, "n_bins": 32, "par_th": 10, "batch_size": 700, "grow_policy": "greedy"}
M = GBTModel(input_dim=f, output_dim=d, tree_struct=ts, optimizers={"algo": "SGD", "lr": 0.2, "start_idx": 0, "stop_idx": d},
             params={"split_score_func": "L2", "generator_type": "Uniform"}, verbose=0, device=DEV)
loss = M.fit(X, y, iterations=5, shuffle=False)
out["s3_loss"] = np.array([loss], np.float64)
out["s3_pred"] = np.asarray(M(X, requires_grad=False, tensor=False)).reshape(n, d)
dump("s3", M.learner._cpp_model)

# ---- S4: ActorCritic model the way GBRL_SB3's PPO drives it: forward with autograd, loss.backward(), step()
n, f, d = 2048, 10, 3
X, y = data(n, f, d, 17)
ts = {"max_depth": 4, "min_data_in_leaf": 0, "n_bins": 64, "par_th": 10, "batch_size": n, "grow_policy": "oblivious"}
AC = ActorCritic(tree_struct=ts, input_dim=f, output_dim=d,
                 policy_optimizer={"policy_algo": "SGD", "policy_lr": 0.1, "start_idx": 0, "stop_idx": d - 1},
                 value_optimizer={"value_algo": "SGD", "value_lr": 0.05, "start_idx": d - 1, "stop_idx": d},
                 shared_tree_struct=True, params={"split_score_func": "Cosine", "generator_type": "Quantile"}, verbose=0, device=DEV)
Xt, yt = T(X), T(y)
for it in range(2):
    pol, val = AC(Xt, requires_grad=True)
    loss = 0.5 * ((pol - yt[:, :d - 1]) ** 2).sum(1).mean() + 0.5 * ((val - yt[:, d - 1]) ** 2).mean()
    loss.backward()
    AC.step()
pol, val = AC(Xt, requires_grad=False)
out["s4_policy"] = pol.detach().cpu().numpy()
out["s4_value"] = val.detach().cpu().numpy()
out["s4_ntrees"] = np.array([AC.get_num_trees()])
dump("s4", AC.learner._cpp_model)

np.savez(OUT, **out)
print("LEARNER_WORKER_OK", ENGINE, backend, flush=True)
os._exit(0)
