"""Host-side mirror of the reference's pybind11 class `gbrl_cpp.GBRL` (gbrl/src/cpp/binding.cpp:421-1134).

Same method names, keyword arguments, argument forms (C-contiguous float32 NumPy array, or the learner's
4-tuple ``(data_ptr, shape, "torch.float32", "cpu"|"cuda")`` from gbrl/common/utils.py:43-60, or ``None``),
shape inference and error behaviour (RuntimeError) for the fit/predict hot path, so that
gbrl/learners/*.py drive it unchanged.  All compute goes through the C-ABI of libgbrl_b200.so; there is
no CPU implementation here.  Out of scope (raise NotImplementedError): categorical features, control
variates, Adam, SHAP, export/plot.
"""
import ctypes as C
import os

import numpy as np

from . import _capi

_SCORE = {"l2": 0, "cosine": 1}
_SCORE_R = {0: "L2", 1: "Cosine"}
_GEN = {"uniform": 0, "quantile": 1}
_GEN_R = {0: "Uniform", 1: "Quantile"}
_GROW = {"greedy": 0, "oblivious": 1}
_GROW_R = {0: "Greedy", 1: "Oblivious"}
_SCHED = {"const": 0, "linear": 1}
_SCHED_R = {0: "Const", 1: "Linear"}


def _enum(table, s, what):
    k = str(s).lower()
    if k not in table:
        raise RuntimeError("Invalid %s: %s" % (what, s))   # types.cpp:31-88 throw on unknown strings
    return table[k]


def _default_ref_threads():
    """omp_get_max_threads() of the reference on this host (utils.h:64-81 partitions its float reductions by it):
    OMP_NUM_THREADS when set, else the core count.  GBRL_B200_REF_THREADS overrides both."""
    for k in ("GBRL_B200_REF_THREADS", "OMP_NUM_THREADS"):
        v = os.environ.get(k, "").split(",")[0].strip()
        if v.isdigit() and int(v) >= 1:
            return int(v)
    return os.cpu_count() or 1


def _torch():
    import torch
    return torch


class _Arg:
    """A borrowed float32 matrix: pointer + shape + host/device flag (binding.cpp:102-199 handle_input_info)."""
    __slots__ = ("ptr", "shape", "dev", "keep")

    def __init__(self, obj, name, func, optional):
        self.keep = None
        if obj is None:
            if not optional:
                raise RuntimeError("%s: %s cannot be None" % (func, name))
            self.ptr, self.shape, self.dev = None, (), 0
            return
        if isinstance(obj, tuple):
            if len(obj) != 4:
                raise RuntimeError("%s: %s tuple must be (data_ptr, shape, dtype, device)" % (func, name))
            ptr, shape, dtype, device = obj
            if "float32" not in str(dtype):
                raise RuntimeError("%s: %s must be float32, got %s" % (func, name, dtype))
            self.ptr, self.shape = int(ptr), tuple(int(s) for s in shape)
            self.dev = 1 if str(device).startswith("cuda") or str(device) == "gpu" else 0
            return
        th = None
        try:
            th = _torch()
        except Exception:   # pragma: no cover
            pass
        if th is not None and isinstance(obj, th.Tensor):
            t = obj.detach()
            if t.dtype != th.float32:
                raise RuntimeError("%s: %s must be float32" % (func, name))
            t = t.contiguous()
            self.keep = t
            self.ptr, self.shape, self.dev = t.data_ptr(), tuple(t.shape), 1 if t.is_cuda else 0
            return
        a = np.asarray(obj)
        if a.dtype != np.float32:
            raise RuntimeError("%s: %s must be a float32 array, got %s" % (func, name, a.dtype))
        if not a.flags["C_CONTIGUOUS"]:
            raise RuntimeError("%s: %s must be C-contiguous" % (func, name))
        self.keep = a
        self.ptr, self.shape, self.dev = a.ctypes.data, tuple(a.shape), 0


class GBRL:
    def __init__(self, input_dim=None, output_dim=None, policy_dim=None, max_depth=4, min_data_in_leaf=0, n_bins=256,
                 par_th=10, cv_beta=0.9, split_score_func="cosine", generator_type="quantile",
                 use_control_variates=False, batch_size=5000, grow_policy="greedy", verbose=0, device="cuda",
                 learner_name="GBRL", ref_threads=None, tie_replay=True, band_kappa=0.0, use_subtraction=True,
                 device_ordinal=None, hist_variant=0, replay_variant=0):
        if isinstance(input_dim, GBRL):                       # copy constructor, binding.cpp:441
            self._init_from(input_dim)
            return
        if use_control_variates:
            raise NotImplementedError("control variates are CPU-only in the reference and out of scope here")
        dev = str(device).lower()
        if dev in ("cpu",):
            raise RuntimeError("gbrl_b200 is a CUDA engine: device='cpu' is not available (no CPU fallback)")
        if not (dev.startswith("cuda") or dev == "gpu"):
            raise RuntimeError("Invalid device: %s" % device)
        if device_ordinal is None:
            device_ordinal = int(dev.split(":")[1]) if ":" in dev else int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("GBRL_B200_USE_LOCAL_RANK") else 0
        self._lib = _capi.lib()
        self._kw = dict(input_dim=int(input_dim), output_dim=int(output_dim),
                        policy_dim=int(policy_dim if policy_dim is not None else output_dim), max_depth=int(max_depth),
                        min_data_in_leaf=int(min_data_in_leaf), n_bins=int(n_bins), par_th=int(par_th),
                        cv_beta=float(cv_beta), split_score_func=split_score_func, generator_type=generator_type,
                        use_control_variates=False, batch_size=int(batch_size), grow_policy=grow_policy,
                        verbose=int(verbose), device="cuda", learner_name=learner_name,
                        ref_threads=int(ref_threads if ref_threads else _default_ref_threads()),
                        tie_replay=bool(tie_replay), band_kappa=float(band_kappa), use_subtraction=bool(use_subtraction),
                        device_ordinal=int(device_ordinal), hist_variant=int(hist_variant), replay_variant=int(replay_variant))
        self._create()

    # ------------------------------------------------------------------ lifetime
    def _create(self):
        k = self._kw
        cfg = _capi.Config(k["input_dim"], k["output_dim"], k["policy_dim"], k["max_depth"], k["min_data_in_leaf"],
                           k["n_bins"], k["par_th"], k["batch_size"], _enum(_SCORE, k["split_score_func"], "split_score_func"),
                           _enum(_GEN, k["generator_type"], "generator_type"), _enum(_GROW, k["grow_policy"], "grow_policy"),
                           k["verbose"], k["device_ordinal"], k["ref_threads"], 1 if k["tie_replay"] else 0,
                           k["band_kappa"], 1 if k["use_subtraction"] else 0, k.get("hist_variant", 0), k.get("replay_variant", 0))
        h = C.c_void_p()
        _capi.check(self._lib.gbrl_b200_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.learner_name = k["learner_name"]
        self._input_dim, self._output_dim, self._max_depth = k["input_dim"], k["output_dim"], k["max_depth"]
        self._oblivious = _enum(_GROW, k["grow_policy"], "grow_policy") == 1

    def _init_from(self, other):
        self._lib = other._lib
        self._kw = dict(other._kw)
        self._create()
        self.set_bias(other.get_bias())
        self.set_feature_weights(other.get_feature_weights())
        fm, num = other.get_feature_mapping()
        self.set_feature_mapping(fm, num, _restore=True)
        for o in other.get_optimizers():
            self.set_optimizer(o["algo"], o["scheduler_func"], o["init_lr"], o["start_idx"], o["stop_idx"], o["stop_lr"], o["T"])
        md = other._meta()
        if md.n_trees > 0:
            e = other.get_ensemble_data()
            self._set_ensemble(e, md.n_num_features)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.gbrl_b200_destroy(h)
            except Exception:   # pragma: no cover
                pass
            self._h = None

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        try:
            th = _torch()
            if th.cuda.is_available():
                return C.c_void_p(th.cuda.current_stream(self._kw["device_ordinal"]).cuda_stream)
        except Exception:   # pragma: no cover
            pass
        return C.c_void_p(0)

    def _meta(self):
        md = _capi.Metadata()
        _capi.check(self._lib.gbrl_b200_get_metadata(self._h, C.byref(md)))
        return md

    def _shapes(self, func, obs, cat, n_samples):
        """binding.cpp:483-523: derive (n_obs_samples, n_num_features) and validate against input_dim."""
        if cat is not None:
            raise NotImplementedError("categorical features are out of scope of the B200 engine (SURVEY 2.1 #19)")
        n_obs, n_num = 0, 0
        if obs.ptr is not None:
            if len(obs.shape) == 1:
                n_num = obs.shape[0] if n_samples == 1 else 1
                n_obs = 1 if n_samples == 1 else obs.shape[0]
            else:
                n_obs, n_num = obs.shape[0], obs.shape[1]
            if n_obs != n_samples:
                raise RuntimeError("Number of observations %d != number of gradient samples %d" % (n_obs, n_samples))
        if n_num != self._input_dim:
            raise RuntimeError("Total number of features %d != correct input dim %d" % (n_num, self._input_dim))
        return n_num

    # ------------------------------------------------------------------ hot path
    def step(self, obs, categorical_obs, grads):
        g = _Arg(grads, "grads", "step", False)
        if len(g.shape) == 1:
            n_samples, gd = (1, g.shape[0]) if self._output_dim > 1 else (g.shape[0], 1)
        else:
            n_samples, gd = g.shape[0], g.shape[1]
        if gd != self._output_dim:
            raise RuntimeError("Gradient output dim %d != correct output dim %d" % (gd, self._output_dim))
        o = _Arg(obs, "obs", "step", True)
        n_num = self._shapes("step", o, categorical_obs, n_samples)
        if o.ptr is None:
            raise RuntimeError("Total number of features 0 != correct input dim %d" % self._input_dim)
        _capi.check(self._lib.gbrl_b200_step(self._h, o.ptr, o.dev, g.ptr, g.dev, n_samples, n_num, self._stream()))

    def fit(self, obs, categorical_obs, targets, iterations, shuffle=True, loss_type="MultiRMSE"):
        if str(loss_type).lower() != "multirmse":
            raise RuntimeError("Invalid loss type: %s" % loss_type)
        t = _Arg(targets, "targets", "fit", False)
        if len(t.shape) == 1:
            n_samples, td = (1, t.shape[0]) if self._output_dim > 1 else (t.shape[0], 1)
        else:
            n_samples, td = t.shape[0], t.shape[1]
        if td != self._output_dim:
            raise RuntimeError("Targets output dim %d != correct output dim %d" % (td, self._output_dim))
        o = _Arg(obs, "obs", "fit", True)
        n_num = self._shapes("fit", o, categorical_obs, n_samples)
        loss = C.c_float(0.0)
        _capi.check(self._lib.gbrl_b200_fit(self._h, o.ptr, o.dev, t.ptr, t.dev, int(iterations), n_samples, n_num,
                                            1 if shuffle else 0, C.byref(loss), self._stream()))
        return float(loss.value)

    # fit() in three pieces (engine extension used by bench.py; see include/gbrl_b200.h)
    def fit_begin(self, obs, targets, shuffle=False):
        t = _Arg(targets, "targets", "fit", False)
        o = _Arg(obs, "obs", "fit", False)
        self._fit_keep = (t, o)
        n_samples = t.shape[0]
        _capi.check(self._lib.gbrl_b200_fit_begin(self._h, o.ptr, o.dev, t.ptr, t.dev, n_samples, o.shape[1], 1 if shuffle else 0, self._stream()))

    def fit_iterate(self, iterations, sync=True):
        _capi.check(self._lib.gbrl_b200_fit_iterate(self._h, int(iterations), 1 if sync else 0, self._stream()))

    def fit_end(self):
        loss = C.c_float(0.0)
        _capi.check(self._lib.gbrl_b200_fit_end(self._h, C.byref(loss), self._stream()))
        self._fit_keep = None
        return float(loss.value)

    PROFILE_CLASSES = ("candidates", "binning", "preprocess", "histogram", "allreduce", "scan", "select_replay",
                       "plan_decide", "partition", "finalize", "predict", "spec_wait")

    def profile(self, enable=True):
        _capi.check(self._lib.gbrl_b200_profile(self._h, 1 if enable else 0))

    def get_profile(self):
        n = len(self.PROFILE_CLASSES)
        ms = (C.c_double * n)()
        ln = (C.c_longlong * n)()
        rows = C.c_longlong(0)
        _capi.check(self._lib.gbrl_b200_get_profile(self._h, ms, ln, n, C.byref(rows)))
        out = {k: {"ms": ms[i], "launches": ln[i]} for i, k in enumerate(self.PROFILE_CLASSES)}
        out["hist_rows"] = rows.value
        return out

    def _predict_shape(self, o, categorical_obs):
        if categorical_obs is not None:
            raise NotImplementedError("categorical features are out of scope of the B200 engine (SURVEY 2.1 #19)")
        if o.ptr is None:
            raise RuntimeError("Cannot call predict without observations!")
        if len(o.shape) == 1:                                  # binding.cpp:846-856
            if o.shape[0] == self._input_dim:
                return 1, o.shape[0]
            return o.shape[0], 1
        return o.shape[0], o.shape[1]

    def predict(self, obs, categorical_obs=None, start_tree_idx=0, stop_tree_idx=0, return_torch=False):
        """Returns a DLPack capsule of a CUDA tensor (what the reference returns for a GPU model,
        binding.cpp:230-261); consume it with torch.from_dlpack like gbrl/learners/gbt_learner.py:487."""
        th = _torch()
        out = self.predict_tensor(obs, categorical_obs, start_tree_idx, stop_tree_idx)
        return th.utils.dlpack.to_dlpack(out)

    def predict_tensor(self, obs, categorical_obs=None, start_tree_idx=0, stop_tree_idx=0):
        th = _torch()
        start = 0 if start_tree_idx is None else int(start_tree_idx)
        stop = 0 if stop_tree_idx is None else int(stop_tree_idx)
        n_trees = self.get_num_trees()
        if start < 0 or (start >= n_trees and n_trees > 0):   # binding.cpp:800-811
            raise RuntimeError("start_tree_idx is out of bounds! Got %d, but valid range is [0, %d]" % (start, n_trees - 1))
        if stop < 0 or stop > n_trees:
            raise RuntimeError("stop_tree_idx is out of bounds! Got %d, but valid range is [0, %d]" % (stop, n_trees))
        o = _Arg(obs, "obs", "predict", True)
        n, f = self._predict_shape(o, categorical_obs)
        if f != self._input_dim:
            raise RuntimeError("Incompatible dataset: received %d features, expected %d" % (f, self._input_dim))
        dev = th.device("cuda", self._kw["device_ordinal"])
        out = th.empty((n, self._output_dim), dtype=th.float32, device=dev)
        _capi.check(self._lib.gbrl_b200_predict(self._h, o.ptr, o.dev, n, f, start, stop, out.data_ptr(), 1, self._stream()))
        return out[:, 0] if self._output_dim == 1 else out                # binding.cpp:282-286

    def predict_numpy(self, obs, start_tree_idx=0, stop_tree_idx=0):
        """Convenience for tests: host result without torch."""
        o = _Arg(obs, "obs", "predict", False)
        n, f = self._predict_shape(o, None)
        out = np.empty((n, self._output_dim), dtype=np.float32)
        _capi.check(self._lib.gbrl_b200_predict(self._h, o.ptr, o.dev, n, f, int(start_tree_idx or 0), int(stop_tree_idx or 0),
                                                out.ctypes.data, 0, self._stream()))
        return out[:, 0] if self._output_dim == 1 else out

    # ------------------------------------------------------------------ setters / getters
    def set_bias(self, bias):
        a = _Arg(bias, "bias", "set_bias", False)
        n = int(np.prod(a.shape)) if a.shape else 1
        _capi.check(self._lib.gbrl_b200_set_bias(self._h, a.ptr, n, a.dev))

    def set_feature_weights(self, w):
        a = _Arg(w, "feature_weights", "set_feature_weights", False)
        n = int(np.prod(a.shape)) if a.shape else 1
        _capi.check(self._lib.gbrl_b200_set_feature_weights(self._h, a.ptr, n, a.dev))

    def set_feature_mapping(self, feature_mapping, mapping_numerics, _restore=False):
        """gbrl.cpp:271-316.  `_restore` (load / copy-constructor): the arrays are put back verbatim -- a model that was only
        ever driven through fit() carries the reference's zero-initialised mapping (all `mapping_numerics` False although
        it has no categorical feature), which is not a request for categorical splits."""
        m = np.ascontiguousarray(feature_mapping, dtype=np.int32)
        n = np.ascontiguousarray(mapping_numerics).astype(np.uint8)
        if m.size != n.size:
            raise RuntimeError("feature_mapping and mapping_numerics must have the same length")
        if not np.all(n) and not _restore:
            raise NotImplementedError("categorical features are out of scope of the B200 engine (SURVEY 2.1 #19)")
        _capi.check(self._lib.gbrl_b200_set_feature_mapping(self._h, m.ctypes.data_as(C.POINTER(C.c_int)),
                                                            n.ctypes.data_as(C.POINTER(C.c_uint8)), int(m.size)))

    def get_bias(self):
        out = np.empty(self._output_dim, np.float32)
        _capi.check(self._lib.gbrl_b200_get_bias(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def get_feature_weights(self):
        out = np.empty(self._input_dim, np.float32)
        _capi.check(self._lib.gbrl_b200_get_feature_weights(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def get_feature_mapping(self):
        m = np.empty(self._input_dim, np.int32)
        n = np.empty(self._input_dim, np.uint8)
        _capi.check(self._lib.gbrl_b200_get_feature_mapping(self._h, m.ctypes.data_as(C.POINTER(C.c_int)),
                                                            n.ctypes.data_as(C.POINTER(C.c_uint8)), None, None))
        return m, n.astype(bool)

    def set_optimizer(self, algo="SGD", scheduler="const", init_lr=1.0, start_idx=0, stop_idx=0, stop_lr=1e-8, T=10000,
                      beta_1=0.9, beta_2=0.999, eps=1e-8, shrinkage=1e-5):
        if str(algo).lower() != "sgd":
            raise RuntimeError("Incompatible GPU optimizer: only SGD is supported on the GPU path (gbrl.cpp:476-481)")
        _capi.check(self._lib.gbrl_b200_set_optimizer(self._h, _enum(_SCHED, scheduler, "scheduler"), float(init_lr),
                                                      int(start_idx), int(stop_idx), float(stop_lr), int(T)))

    def get_optimizers(self):
        out = []
        for i in range(self._lib.gbrl_b200_n_optimizers(self._h)):
            s, a, b, T = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            lr, slr = C.c_float(), C.c_float()
            _capi.check(self._lib.gbrl_b200_get_optimizer(self._h, i, C.byref(s), C.byref(lr), C.byref(a), C.byref(b), C.byref(slr), C.byref(T)))
            out.append({"algo": "SGD", "init_lr": lr.value, "start_idx": a.value, "stop_idx": b.value,
                        "scheduler_func": _SCHED_R[s.value], "stop_lr": slr.value, "T": T.value,
                        "beta_1": 0.9, "beta_2": 0.999, "eps]": 1e-8})
        return out

    def get_scheduler_lrs(self):
        n = self._lib.gbrl_b200_n_optimizers(self._h)
        out = np.zeros(max(n, 1), np.float32)
        _capi.check(self._lib.gbrl_b200_get_scheduler_lrs(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out[:n]

    def get_num_trees(self):
        return int(self._meta().n_trees)

    def get_iteration(self):
        return int(self._meta().iteration)

    def get_device(self):
        return "cuda"

    def get_learner_name(self):
        return self.learner_name

    def to_device(self, device):
        if not str(device).lower().startswith(("cuda", "gpu")):
            raise RuntimeError("gbrl_b200 models live on the GPU; to_device('%s') is not available" % device)

    def get_metadata(self):
        md = self._meta()
        return {"input_dim": md.input_dim, "output_dim": md.output_dim, "policy_dim": md.policy_dim,
                "split_score_func": _SCORE_R[md.split_score_func], "generator_type": _GEN_R[md.generator_type],
                "use_control_variates": False, "verbose": md.verbose, "max_depth": md.max_depth,
                "min_data_in_leaf": md.min_data_in_leaf, "n_bins": md.n_bins, "par_th": md.par_th,
                "batch_size": md.batch_size, "grow_policy": _GROW_R[md.grow_policy], "iteration": md.iteration}

    def get_stats(self):
        md = self._meta()
        return {"kernel_launches": md.kernel_launches, "replay_items": md.replay_items, "replay_nodes": md.replay_nodes,
                "replay_overflow": md.replay_overflow, "nodes_evaluated": md.nodes_evaluated, "max_noise_ratio": float(md.max_noise_ratio),
                "n_trees": md.n_trees, "n_leaves": md.n_leaves, "chain_blocks_fast": md.chain_blocks_fast,
                "chain_blocks_slow": md.chain_blocks_slow, "chain_lanes_seq": md.chain_lanes_seq, "replay_flips": md.replay_flips,
                "spec_trees": md.spec_trees, "spec_rollbacks": md.spec_rollbacks}

    def get_ensemble_data(self):
        """binding.cpp:330-390: dict of owning NumPy arrays in the reference layout."""
        md = self._meta()
        nt, nl, d, D = md.n_trees, md.n_leaves, self._max_depth, self._output_dim
        S = nt if self._oblivious else nl
        e = {"tree_indices": np.zeros(nt, np.int32), "depths": np.zeros(S, np.int32),
             "values": np.zeros((nl, D), np.float32), "feature_indices": np.zeros((S, d), np.int32),
             "feature_values": np.zeros((S, d), np.float32), "edge_weights": np.zeros((nl, d), np.float32),
             "inequality_directions": np.zeros((nl, d), np.uint8)}
        ip, fp, u8p = C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        _capi.check(self._lib.gbrl_b200_get_ensemble(
            self._h, e["tree_indices"].ctypes.data_as(ip), e["depths"].ctypes.data_as(ip), e["values"].ctypes.data_as(fp),
            e["feature_indices"].ctypes.data_as(ip), e["feature_values"].ctypes.data_as(fp),
            e["edge_weights"].ctypes.data_as(fp), e["inequality_directions"].ctypes.data_as(u8p)))
        e["inequality_directions"] = e["inequality_directions"].astype(bool)
        e["is_numerics"] = np.ones((S, d), bool)
        e["bias"] = self.get_bias()
        e["feature_weights"] = self.get_feature_weights()
        fm, num = self.get_feature_mapping()
        e["feature_mapping"], e["mapping_numerics"] = fm, num
        return e

    def _set_ensemble(self, e, n_num_features):
        ip, fp, u8p = C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        ti = np.ascontiguousarray(e["tree_indices"], np.int32)
        dp = np.ascontiguousarray(e["depths"], np.int32)
        va = np.ascontiguousarray(e["values"], np.float32)
        fi = np.ascontiguousarray(e["feature_indices"], np.int32)
        fv = np.ascontiguousarray(e["feature_values"], np.float32)
        ew = np.ascontiguousarray(e["edge_weights"], np.float32)
        iq = np.ascontiguousarray(e["inequality_directions"]).astype(np.uint8)
        nt = int(ti.size)
        nl = int(va.shape[0]) if va.ndim == 2 else int(va.size // self._output_dim)
        _capi.check(self._lib.gbrl_b200_set_ensemble(
            self._h, nt, nl, ti.ctypes.data_as(ip), dp.ctypes.data_as(ip), va.ctypes.data_as(fp), fi.ctypes.data_as(ip),
            fv.ctypes.data_as(fp), ew.ctypes.data_as(fp), iq.ctypes.data_as(u8p), int(n_num_features)))

    def get_candidates(self):
        n = C.c_int(0)
        out = np.empty(self._input_dim * self._kw["n_bins"], np.float32)
        _capi.check(self._lib.gbrl_b200_get_candidates(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n)))
        return out[:n.value]

    def get_root_scores(self):
        n = C.c_int(0)
        out = np.empty(self._input_dim * self._kw["n_bins"], np.float32)
        _capi.check(self._lib.gbrl_b200_get_root_scores(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n)))
        return out[:n.value]

    # ------------------------------------------------------------------ multi-GPU
    def init_distributed(self, rank=None, world_size=None):
        """Joins this model to the job's torch.distributed group: rank 0 creates the NCCL id, it is broadcast
        through torch.distributed, and the engine opens its own communicator for the per-level all-reduce."""
        import torch.distributed as dist
        th = _torch()
        rank = dist.get_rank() if rank is None else rank
        world_size = dist.get_world_size() if world_size is None else world_size
        buf = np.zeros(128, np.uint8)
        if rank == 0:
            _capi.check(self._lib.gbrl_b200_dist_unique_id(buf.ctypes.data_as(C.POINTER(C.c_uint8))))
        t = th.from_numpy(buf)
        if dist.get_backend() == "nccl":
            t = t.cuda(self._kw["device_ordinal"])
        dist.broadcast(t, 0)
        buf = t.cpu().numpy().copy()
        _capi.check(self._lib.gbrl_b200_dist_init(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint8)), int(rank), int(world_size)))

    # ------------------------------------------------------------------ statics / out of scope
    @staticmethod
    def cuda_available():
        return bool(_capi.lib().gbrl_b200_cuda_available())

    def _nyi(self, *a, **k):
        raise NotImplementedError("not on the fit/predict hot path (SURVEY.md 8): use the reference implementation")

    export = print_tree = plot_tree = print_ensemble_metadata = tree_shap = ensemble_shap = _nyi

    # ------------------------------------------------------------------ checkpoint (reference wire format)
    def save(self, path):
        """GBRL::saveToFile (gbrl.cpp:1130-1173): writes the reference's `.gbrl_model` format, so a model trained here
        loads in the reference (CPU) and vice versa."""
        from . import model_io
        md = self._meta()
        meta = {"n_leaves": md.n_leaves, "n_trees": md.n_trees, "input_dim": md.input_dim, "output_dim": md.output_dim,
                "policy_dim": md.policy_dim, "max_depth": md.max_depth, "min_data_in_leaf": md.min_data_in_leaf,
                "n_bins": md.n_bins, "par_th": md.par_th, "cv_beta": self._kw["cv_beta"], "verbose": md.verbose,
                "batch_size": md.batch_size, "use_cv": 0, "split_score_func": md.split_score_func,
                "generator_type": md.generator_type, "grow_policy": md.grow_policy, "n_num_features": md.n_num_features,
                "n_cat_features": 0, "iteration": md.iteration}
        e = self.get_ensemble_data()
        rev_num = np.empty(self._input_dim, np.int32)
        rev_cat = np.empty(self._input_dim, np.int32)
        _capi.check(self._lib.gbrl_b200_get_feature_mapping(self._h, None, None, rev_num.ctypes.data_as(C.POINTER(C.c_int)),
                                                            rev_cat.ctypes.data_as(C.POINTER(C.c_int))))
        e["reverse_num_feature_mapping"], e["reverse_cat_feature_mapping"] = rev_num, rev_cat
        model_io.write_model(path, meta, e, self.get_optimizers(), self.learner_name)
        return 0

    @staticmethod
    def load(path, device="cuda", **engine_kwargs):
        """GBRL::loadFromFile (gbrl.cpp:1175-1252) for numerical-feature SGD models; the model lands on the GPU."""
        from . import model_io
        meta, e, opts, name = model_io.read_model(path)
        for k in ("bias", "feature_weights", "feature_mapping", "mapping_numerics"):
            if e[k] is None:
                raise RuntimeError("corrupt model file: section %s is absent" % k)
        if meta["n_trees"] > 0:
            for k in ("tree_indices", "depths", "values", "feature_indices", "feature_values", "edge_weights", "inequality_directions"):
                if e[k] is None:
                    raise RuntimeError("corrupt model file: section %s is absent" % k)
        categorical = meta["n_cat_features"] != 0
        if not categorical and meta["n_trees"] > 0 and e["is_numerics"] is not None and meta["max_depth"] > 0:
            # only the first depths[r] entries of a row are splits; the unused tail is zero-initialised (types.cpp:232-260)
            used = np.arange(meta["max_depth"])[None, :] < np.asarray(e["depths"])[:, None]
            categorical = bool(np.any(used & ~e["is_numerics"]))
        if categorical:
            raise NotImplementedError("categorical features are out of scope of the B200 engine (SURVEY 2.1 #19)")
        m = GBRL(input_dim=meta["input_dim"], output_dim=meta["output_dim"], policy_dim=meta["policy_dim"],
                 max_depth=meta["max_depth"], min_data_in_leaf=meta["min_data_in_leaf"], n_bins=meta["n_bins"],
                 par_th=meta["par_th"], cv_beta=meta["cv_beta"], split_score_func=_SCORE_R[meta["split_score_func"]],
                 generator_type=_GEN_R[meta["generator_type"]], batch_size=meta["batch_size"],
                 grow_policy=_GROW_R[meta["grow_policy"]], verbose=meta["verbose"], device=device, learner_name=name,
                 **engine_kwargs)
        m.set_bias(e["bias"])
        m.set_feature_weights(e["feature_weights"])
        m.set_feature_mapping(e["feature_mapping"], e["mapping_numerics"], _restore=True)
        for o in opts:
            m.set_optimizer(o["algo"], o["scheduler_func"], o["init_lr"], o["start_idx"], o["stop_idx"], o["stop_lr"], o["T"])
        if meta["n_trees"] > 0:
            m._set_ensemble(e, meta["n_num_features"])
        _capi.check(m._lib.gbrl_b200_set_iteration(m._h, int(meta["iteration"])))
        return m
