"""ctypes binding of include/gbrl_b200.h (libgbrl_b200.so).

The library is loaded eagerly and loudly: there is no CPU fallback and no alternative implementation.
If the shared object is missing this module raises ImportError telling how to build it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GBRL_B200_LIB") or os.path.join(_HERE, "lib", "libgbrl_b200.so")   # override: A/B runs of two builds


class Config(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "input_dim", "output_dim", "policy_dim", "max_depth", "min_data_in_leaf", "n_bins", "par_th",
        "batch_size", "split_score_func", "generator_type", "grow_policy", "verbose", "device_ordinal",
        "ref_threads", "tie_replay")] + [("band_kappa", C.c_float), ("use_subtraction", C.c_int), ("hist_variant", C.c_int), ("replay_variant", C.c_int)]


class Metadata(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "input_dim", "output_dim", "policy_dim", "max_depth", "min_data_in_leaf", "n_bins", "par_th",
        "batch_size", "split_score_func", "generator_type", "grow_policy", "verbose", "n_num_features",
        "n_cat_features", "n_trees", "n_leaves", "iteration")] + [(n, C.c_longlong) for n in (
            "kernel_launches", "replay_items", "replay_nodes", "replay_overflow", "nodes_evaluated")] + [("max_noise_ratio", C.c_float)] + [
                (n, C.c_longlong) for n in ("chain_blocks_fast", "chain_blocks_slow", "chain_lanes_seq", "replay_flips", "spec_trees", "spec_rollbacks")]


# every symbol include/gbrl_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "gbrl_b200_last_error", "gbrl_b200_cuda_available", "gbrl_b200_create", "gbrl_b200_destroy",
    "gbrl_b200_set_bias", "gbrl_b200_set_feature_weights", "gbrl_b200_set_feature_mapping", "gbrl_b200_get_bias",
    "gbrl_b200_get_feature_weights", "gbrl_b200_get_feature_mapping", "gbrl_b200_set_optimizer",
    "gbrl_b200_n_optimizers", "gbrl_b200_get_optimizer", "gbrl_b200_get_scheduler_lrs", "gbrl_b200_step",
    "gbrl_b200_fit", "gbrl_b200_fit_begin", "gbrl_b200_fit_iterate", "gbrl_b200_fit_end", "gbrl_b200_profile",
    "gbrl_b200_get_profile", "gbrl_b200_predict", "gbrl_b200_get_metadata", "gbrl_b200_get_ensemble",
    "gbrl_b200_set_ensemble", "gbrl_b200_set_iteration", "gbrl_b200_get_candidates", "gbrl_b200_get_root_scores", "gbrl_b200_dist_unique_id",
    "gbrl_b200_dist_init", "gbrl_b200_dist_shutdown", "gbrl_b200_microbench", "gbrl_b200_diag_chain_sums",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "gbrl_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C gbrl_b200`). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, fp, ip, u8p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint8)
    L.gbrl_b200_last_error.restype = C.c_char_p
    L.gbrl_b200_cuda_available.restype = C.c_int
    L.gbrl_b200_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.gbrl_b200_destroy.argtypes = [vp]
    L.gbrl_b200_destroy.restype = None
    L.gbrl_b200_set_bias.argtypes = [vp, vp, C.c_int, C.c_int]
    L.gbrl_b200_set_feature_weights.argtypes = [vp, vp, C.c_int, C.c_int]
    L.gbrl_b200_set_feature_mapping.argtypes = [vp, ip, u8p, C.c_int]
    L.gbrl_b200_get_bias.argtypes = [vp, fp]
    L.gbrl_b200_get_feature_weights.argtypes = [vp, fp]
    L.gbrl_b200_get_feature_mapping.argtypes = [vp, ip, u8p, ip, ip]
    L.gbrl_b200_set_optimizer.argtypes = [vp, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int]
    L.gbrl_b200_n_optimizers.argtypes = [vp]
    L.gbrl_b200_get_optimizer.argtypes = [vp, C.c_int, ip, fp, ip, ip, fp, ip]
    L.gbrl_b200_get_scheduler_lrs.argtypes = [vp, fp]
    L.gbrl_b200_step.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp]
    L.gbrl_b200_fit.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, vp]
    L.gbrl_b200_predict.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]
    L.gbrl_b200_fit_begin.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.gbrl_b200_fit_iterate.argtypes = [vp, C.c_int, C.c_int, vp]
    L.gbrl_b200_fit_end.argtypes = [vp, fp, vp]
    L.gbrl_b200_profile.argtypes = [vp, C.c_int]
    L.gbrl_b200_get_profile.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_longlong)]
    L.gbrl_b200_get_metadata.argtypes = [vp, C.POINTER(Metadata)]
    L.gbrl_b200_get_ensemble.argtypes = [vp, ip, ip, fp, ip, fp, fp, u8p]
    L.gbrl_b200_set_ensemble.argtypes = [vp, C.c_int, C.c_int, ip, ip, fp, ip, fp, fp, u8p, C.c_int]
    L.gbrl_b200_set_iteration.argtypes = [vp, C.c_int]
    L.gbrl_b200_get_candidates.argtypes = [vp, fp, ip]
    L.gbrl_b200_get_root_scores.argtypes = [vp, fp, ip]
    L.gbrl_b200_dist_unique_id.argtypes = [u8p]
    L.gbrl_b200_dist_init.argtypes = [vp, u8p, C.c_int, C.c_int]
    L.gbrl_b200_dist_shutdown.argtypes = [vp]
    L.gbrl_b200_microbench.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.gbrl_b200_diag_chain_sums.argtypes = [fp, C.c_longlong, C.c_int, C.c_int, C.c_int, fp, fp, fp, C.c_int, C.POINTER(C.c_double)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError(lib().gbrl_b200_last_error().decode("utf-8", "replace"))
