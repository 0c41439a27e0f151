"""`.gbrl_model` wire format of the reference, read and written without the reference.

Layout restated from gbrl/src/cpp (v1.1.6), little-endian, natural struct padding of x86-64:
  serializationHeader     gbrl.cpp:1138 / types.h:312-318   u16 major, minor, patch; pad; u64 0; u32 0; pad   (24 B)
  ensembleMetaData        gbrl.cpp:1139 / types.h:218-242   raw struct                                          (80 B)
  u8 parallel_predict, u8 use_cv                            gbrl.cpp:1140-1144
  u64 name_length, name bytes                               gbrl.cpp:1147-1149
  ensemble arrays         types.cpp:681-767                 each: u8 NULL_CHECK (1 = present) + raw data, trimmed to n_trees / n_leaves
  i32 n_optimizers, then per optimizer                      gbrl.cpp:1153-1164, optimizer.cpp:120-131, scheduler.cpp:64-108
      u8 algo (0 SGD), i32 start_idx, i32 stop_idx, u8 scheduler (0 Const, 1 Linear), f32 init_lr [, f32 stop_lr, i32 T]
Only numerical features and SGD are supported (the scope of this engine); categorical arrays are written empty-valued.
"""
import struct

import numpy as np

VERSION = (1, 1, 6)
MAX_CHAR_SIZE = 128
_META_FMT = "<13i f 2i 4B 3i"      # 80 bytes, see ensembleMetaData
_META_KEYS = ("n_leaves", "n_trees", "max_trees", "max_leaves", "max_trees_batch", "max_leaves_batch", "input_dim",
              "output_dim", "policy_dim", "max_depth", "min_data_in_leaf", "n_bins", "par_th", "cv_beta", "verbose",
              "batch_size", "use_cv", "split_score_func", "generator_type", "grow_policy", "n_num_features",
              "n_cat_features", "iteration")
assert struct.calcsize(_META_FMT) == 80


def _arr(f, a, dtype):
    f.write(b"\x01")
    f.write(np.ascontiguousarray(a, dtype=dtype).tobytes())


def write_model(path, meta, ens, optimizers, learner_name="GBRL"):
    """meta: dict with _META_KEYS (enums as ints: score 0 L2 / 1 Cosine, generator 0 Uniform / 1 Quantile, grow 0 greedy /
    1 oblivious); ens: dict in the layout of get_ensemble_data(); optimizers: list of dicts (scheduler_func 'Const'|'Linear')."""
    md, D, I = meta["max_depth"], meta["output_dim"], meta["input_dim"]
    nt, nl = meta["n_trees"], meta["n_leaves"]
    S = nt if meta["grow_policy"] == 1 else nl
    m = dict(meta)
    m.setdefault("max_trees", max(nt + 1024, 2048))
    m.setdefault("max_leaves", m["max_trees"] << md)
    m.setdefault("max_trees_batch", 25000)
    m.setdefault("max_leaves_batch", 25000 << md)
    m.setdefault("n_cat_features", 0)
    m.setdefault("cv_beta", 0.9)
    with open(path, "wb") as f:
        f.write(struct.pack("<3H2xQI4x", *VERSION, 0, 0))
        f.write(struct.pack(_META_FMT, *[m[k] if k != "use_cv" else int(bool(m[k])) for k in _META_KEYS]))
        f.write(struct.pack("<BB", 1, 0))                         # parallel_predict (SGD), use_cv
        name = learner_name.encode()
        f.write(struct.pack("<Q", len(name))); f.write(name)
        _arr(f, ens["bias"], np.float32)
        _arr(f, ens["feature_weights"], np.float32)
        _arr(f, ens["tree_indices"][:nt], np.int32)
        _arr(f, ens["depths"][:S], np.int32)
        _arr(f, np.asarray(ens["values"]).reshape(nl, D), np.float32)
        _arr(f, np.asarray(ens["feature_indices"]).reshape(S, md), np.int32)
        _arr(f, np.asarray(ens["feature_values"]).reshape(S, md), np.float32)
        _arr(f, np.asarray(ens["edge_weights"]).reshape(nl, md), np.float32)
        _arr(f, ens["reverse_num_feature_mapping"], np.int32)
        _arr(f, ens["reverse_cat_feature_mapping"], np.int32)
        _arr(f, ens["feature_mapping"], np.int32)
        _arr(f, np.asarray(ens["mapping_numerics"]).astype(np.uint8), np.uint8)
        _arr(f, np.ones((S, md), np.uint8), np.uint8)             # is_numerics
        _arr(f, np.asarray(ens["inequality_directions"]).reshape(nl, md).astype(np.uint8), np.uint8)
        _arr(f, np.zeros((S, md, MAX_CHAR_SIZE), np.uint8), np.uint8)   # categorical_values
        f.write(struct.pack("<i", len(optimizers)))
        for o in optimizers:
            if str(o.get("algo", "SGD")).upper() != "SGD":
                raise NotImplementedError("only SGD optimizers are serialised by this engine")
            f.write(struct.pack("<Bii", 0, int(o["start_idx"]), int(o["stop_idx"])))
            if str(o["scheduler_func"]).lower() == "const":
                f.write(struct.pack("<Bf", 0, float(o["init_lr"])))
            else:
                f.write(struct.pack("<Bffi", 1, float(o["init_lr"]), float(o["stop_lr"]), int(o["T"])))


def read_model(path):
    """Returns (meta dict, ensemble dict, optimizers list, learner_name)."""
    with open(path, "rb") as f:
        buf = f.read()
    off = 0

    def take(fmt):
        nonlocal off
        v = struct.unpack_from(fmt, buf, off)
        off += struct.calcsize(fmt)
        return v

    major, minor, patch, _, _ = take("<3H2xQI4x")
    meta = dict(zip(_META_KEYS, take(_META_FMT)))
    meta["version"] = (major, minor, patch)
    take("<BB")
    (nlen,) = take("<Q")
    name = buf[off:off + nlen].decode(errors="replace"); off += nlen
    md, D, I = meta["max_depth"], meta["output_dim"], meta["input_dim"]
    nt, nl = meta["n_trees"], meta["n_leaves"]
    S = nt if meta["grow_policy"] == 1 else nl

    def arr(count, dtype, shape=None):
        nonlocal off
        if off >= len(buf):
            raise RuntimeError("truncated .gbrl_model file (offset %d of %d bytes)" % (off, len(buf)))
        present = buf[off]; off += 1
        if not present:
            return None
        if off + count * np.dtype(dtype).itemsize > len(buf):
            raise RuntimeError("truncated .gbrl_model file: section of %d x %s at offset %d exceeds %d bytes" % (
                count, np.dtype(dtype).name, off, len(buf)))
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=off).copy()
        off += a.nbytes
        return a.reshape(shape) if shape is not None else a

    ens = {}
    ens["bias"] = arr(D, np.float32)
    ens["feature_weights"] = arr(I, np.float32)
    ens["tree_indices"] = arr(nt, np.int32)
    ens["depths"] = arr(S, np.int32)
    ens["values"] = arr(nl * D, np.float32, (nl, D))
    ens["feature_indices"] = arr(S * md, np.int32, (S, md))
    ens["feature_values"] = arr(S * md, np.float32, (S, md))
    ens["edge_weights"] = arr(nl * md, np.float32, (nl, md))
    ens["reverse_num_feature_mapping"] = arr(I, np.int32)
    ens["reverse_cat_feature_mapping"] = arr(I, np.int32)
    ens["feature_mapping"] = arr(I, np.int32)
    mn = arr(I, np.uint8)
    ens["mapping_numerics"] = None if mn is None else mn.astype(bool)
    isn = arr(S * md, np.uint8, (S, md))
    ens["is_numerics"] = None if isn is None else isn.astype(bool)
    iq = arr(nl * md, np.uint8, (nl, md))
    ens["inequality_directions"] = None if iq is None else iq.astype(bool)
    ens["categorical_values"] = arr(S * md * MAX_CHAR_SIZE, np.uint8)
    (n_opts,) = take("<i")
    opts = []
    for _ in range(n_opts):
        algo, start, stop = take("<Bii")
        (sched,) = take("<B")
        if algo != 0:
            raise NotImplementedError("Adam optimizers are out of scope of this engine")
        if sched == 0:
            (lr,) = take("<f")
            opts.append({"algo": "SGD", "scheduler_func": "Const", "init_lr": lr, "start_idx": start, "stop_idx": stop,
                         "stop_lr": 1e-8, "T": 10000})
        else:
            lr, slr, T = take("<ffi")
            opts.append({"algo": "SGD", "scheduler_func": "Linear", "init_lr": lr, "start_idx": start, "stop_idx": stop,
                         "stop_lr": slr, "T": T})
    return meta, ens, opts, name
