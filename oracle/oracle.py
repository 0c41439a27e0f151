"""ctypes front-end of the plain-C oracle (oracle/gbrl_oracle.c) and loader of the compiled reference.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg.  Nothing under gbrl_b200/ imports this module.

`Oracle` mirrors the subset of the reference's `gbrl_cpp.GBRL` surface that is on the hot path
(binding.cpp:421-1134): step / fit / predict / set_bias / set_feature_weights / set_feature_mapping /
set_optimizer / get_ensemble_data.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SCORE = {"l2": 0, "cosine": 1}
GEN = {"uniform": 0, "quantile": 1}
GROW = {"greedy": 0, "oblivious": 1}
SCHED = {"const": 0, "linear": 1}


def build(verbose=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is mounted)."""
    out = subprocess.run(["make", "-s", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        fp, ip, u8p = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint8)
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_create.argtypes = [C.c_int] * 11
        lib.oracle_destroy.argtypes = [C.c_void_p]
        lib.oracle_set_bias.argtypes = [C.c_void_p, fp]
        lib.oracle_set_feature_weights.argtypes = [C.c_void_p, fp]
        lib.oracle_set_feature_mapping.argtypes = [C.c_void_p, ip, u8p]
        lib.oracle_set_optimizer.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int]
        lib.oracle_set_optimizer.restype = C.c_int
        for name in ("oracle_n_trees", "oracle_n_leaves", "oracle_iteration"):
            getattr(lib, name).argtypes = [C.c_void_p]
            getattr(lib, name).restype = C.c_int
        lib.oracle_get_ensemble.argtypes = [C.c_void_p, ip, ip, fp, ip, fp, fp, u8p]
        lib.oracle_step.argtypes = [C.c_void_p, fp, fp, C.c_int, C.c_int]
        lib.oracle_step.restype = C.c_int
        lib.oracle_predict.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp]
        lib.oracle_predict.restype = C.c_int
        lib.oracle_fit.argtypes = [C.c_void_p, fp, fp, C.c_int, C.c_int, C.c_int]
        lib.oracle_fit.restype = C.c_float
        lib.oracle_build_grads.argtypes = [C.c_void_p, fp, C.c_int, fp]
        lib.oracle_root_scores.argtypes = [C.c_void_p, fp, fp, C.c_int, C.c_int, fp, fp]
        lib.oracle_root_scores.restype = C.c_int
        _LIB = lib
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Oracle:
    def __init__(self, input_dim, output_dim, max_depth=4, min_data_in_leaf=0, n_bins=256, par_th=10,
                 split_score_func="cosine", generator_type="quantile", batch_size=5000,
                 grow_policy="greedy", ref_threads=1, **_ignored):
        self.lib = _lib()
        self.input_dim, self.output_dim, self.max_depth = input_dim, output_dim, max_depth
        self.grow_policy = grow_policy.lower()
        self.h = self.lib.oracle_create(input_dim, output_dim, max_depth, min_data_in_leaf, n_bins, par_th,
                                        batch_size, SCORE[split_score_func.lower()],
                                        GEN[generator_type.lower()], GROW[self.grow_policy], ref_threads)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_destroy(self.h)
            self.h = None

    def set_bias(self, b):
        b = _f32(b); assert b.size == self.output_dim
        self.lib.oracle_set_bias(self.h, _fp(b))

    def set_feature_weights(self, w):
        w = _f32(w); assert w.size == self.input_dim
        self.lib.oracle_set_feature_weights(self.h, _fp(w))

    def set_feature_mapping(self, mapping, numerics):
        m = np.ascontiguousarray(mapping, dtype=np.int32)
        n = np.ascontiguousarray(numerics, dtype=np.uint8)
        self.lib.oracle_set_feature_mapping(self.h, m.ctypes.data_as(C.POINTER(C.c_int)), n.ctypes.data_as(C.POINTER(C.c_uint8)))

    def set_optimizer(self, algo="SGD", scheduler="const", init_lr=1.0, start_idx=0, stop_idx=0, stop_lr=1e-8, T=10000, **_):
        assert algo.upper() == "SGD"
        rc = self.lib.oracle_set_optimizer(self.h, SCHED[scheduler.lower()], init_lr, start_idx, stop_idx, stop_lr, T)
        if rc != 0:
            raise RuntimeError("invalid optimizer (%d)" % rc)

    def get_num_trees(self):
        return self.lib.oracle_n_trees(self.h)

    def step(self, obs, grads):
        obs, grads = _f32(obs), _f32(grads)
        n, f = obs.shape
        assert f == self.input_dim and grads.size == n * self.output_dim
        self.lib.oracle_step(self.h, _fp(obs), _fp(grads), n, f)

    def fit(self, obs, targets, iterations):
        obs, targets = _f32(obs), _f32(targets)
        n, f = obs.shape
        return float(self.lib.oracle_fit(self.h, _fp(obs), _fp(targets), iterations, n, f))

    def predict(self, obs, start_tree_idx=0, stop_tree_idx=0):
        obs = _f32(obs)
        n, f = obs.shape
        preds = np.zeros((n, self.output_dim), dtype=np.float32)
        rc = self.lib.oracle_predict(self.h, _fp(obs), n, f, start_tree_idx, stop_tree_idx, _fp(preds))
        if rc != 0:
            raise RuntimeError("predict failed (%d)" % rc)
        return preds[:, 0] if self.output_dim == 1 else preds

    def build_grads(self, grads):
        grads = _f32(grads)
        out = np.empty_like(grads)
        self.lib.oracle_build_grads(self.h, _fp(grads), grads.shape[0], _fp(out))
        return out

    def root_scores(self, obs, grads):
        obs, grads = _f32(obs), _f32(grads)
        n, f = obs.shape
        nb = self.lib.oracle_root_scores  # noqa
        scores = np.empty(f * 65536, dtype=np.float32)
        thr = np.empty(f * 65536, dtype=np.float32)
        nc = self.lib.oracle_root_scores(self.h, _fp(obs), _fp(grads), n, f, _fp(scores), _fp(thr))
        return scores[:nc].copy(), thr[:nc].copy()

    def get_ensemble_data(self):
        nt, nl = self.lib.oracle_n_trees(self.h), self.lib.oracle_n_leaves(self.h)
        d, D = self.max_depth, self.output_dim
        S = nt if self.grow_policy == "oblivious" else nl
        out = {
            "tree_indices": np.zeros(nt, np.int32), "depths": np.zeros(S, np.int32),
            "values": np.zeros((nl, D), np.float32), "feature_indices": np.zeros((S, d), np.int32),
            "feature_values": np.zeros((S, d), np.float32), "edge_weights": np.zeros((nl, d), np.float32),
            "inequality_directions": np.zeros((nl, d), np.uint8),
        }
        ip, u8p = C.POINTER(C.c_int), C.POINTER(C.c_uint8)
        self.lib.oracle_get_ensemble(self.h, out["tree_indices"].ctypes.data_as(ip), out["depths"].ctypes.data_as(ip),
                                     _fp(out["values"]), out["feature_indices"].ctypes.data_as(ip),
                                     _fp(out["feature_values"]), _fp(out["edge_weights"]),
                                     out["inequality_directions"].ctypes.data_as(u8p))
        out["inequality_directions"] = out["inequality_directions"].astype(bool)
        return out


# ----------------------------------------------------------------------------------------------------
# the compiled reference (oracle/_ref): returns the pybind11 module `gbrl_cpp` or None
def load_reference():
    ref_dir = os.path.join(_HERE, "_ref")
    if not os.path.isdir(ref_dir) or not any(f.startswith("gbrl_cpp") and f.endswith(".so") for f in os.listdir(ref_dir)):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import gbrl_cpp  # noqa
        return gbrl_cpp
    except Exception:  # pragma: no cover - e.g. ABI mismatch on a foreign box
        return None


def make_reference(gbrl_cpp, input_dim, output_dim, max_depth=4, min_data_in_leaf=0, n_bins=256, par_th=10,
                   split_score_func="cosine", generator_type="quantile", batch_size=5000,
                   grow_policy="greedy", lrs=None, identity_mapping=True, feature_weights=None, bias=None):
    """Construct a reference GBRL on CPU configured the way gbrl/learners/gbt_learner.py does."""
    m = gbrl_cpp.GBRL(input_dim=input_dim, output_dim=output_dim, policy_dim=output_dim, max_depth=max_depth,
                      min_data_in_leaf=min_data_in_leaf, n_bins=n_bins, par_th=par_th, cv_beta=0.9,
                      split_score_func=split_score_func, generator_type=generator_type,
                      use_control_variates=False, batch_size=batch_size, grow_policy=grow_policy, verbose=0,
                      device="cpu")
    m.set_bias(np.zeros(output_dim, np.float32) if bias is None else _f32(bias))
    m.set_feature_weights(np.ones(input_dim, np.float32) if feature_weights is None else _f32(feature_weights))
    if identity_mapping:
        m.set_feature_mapping(np.arange(input_dim, dtype=np.int32), np.ones(input_dim, dtype=bool))
    for (lr, a, b) in (lrs or [(0.1, 0, output_dim)]):
        m.set_optimizer("SGD", "const", lr, a, b)
    return m
