"""CPU suite: the `.gbrl_model` wire format against the compiled reference, both directions."""
import os

import numpy as np
import pytest

from helpers import synth

KEYS = ("tree_indices", "depths", "values", "feature_indices", "feature_values", "edge_weights", "inequality_directions",
        "bias", "feature_weights", "feature_mapping", "mapping_numerics")


def _train_reference(grow, tmp_path):
    from oracle.oracle import load_reference, make_reference
    ref = load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    n, f, d = 1500, 7, 2
    X, y = synth(n, f, d, 13)
    m = make_reference(ref, input_dim=f, output_dim=d, max_depth=4, n_bins=32, split_score_func="cosine", generator_type="quantile",
                       batch_size=n, grow_policy=grow, lrs=[(0.1, 0, 1), (0.05, 1, 2)], bias=[0.5, -1.0],
                       feature_weights=1.0 + 0.1 * np.arange(f))
    for it in range(3):
        p = np.array(m.predict(X, None), copy=True).reshape(n, d)
        m.step(X, None, (p - y).astype(np.float32))
    return ref, m, X


@pytest.mark.parametrize("grow", ["greedy", "oblivious"])
def test_read_reference_file_and_write_it_back(grow, tmp_path):
    from gbrl_b200 import model_io
    ref, m, X = _train_reference(grow, tmp_path)
    path = str(tmp_path / "ref.gbrl_model")
    assert m.save(path) == 0
    meta, e, opts, name = model_io.read_model(path)
    ens = m.get_ensemble_data()                      # once per model (reference quirk)
    assert meta["n_trees"] == 3 and meta["grow_policy"] == (1 if grow == "oblivious" else 0) and meta["version"] == (1, 1, 6)
    for k in KEYS:
        assert np.array_equal(np.asarray(ens[k]).reshape(np.asarray(e[k]).shape), e[k]), k
    assert [(o["start_idx"], o["stop_idx"], round(o["init_lr"], 6)) for o in opts] == [(0, 1, 0.1), (1, 2, 0.05)]
    # our writer -> the reference's loader: identical predictions
    path2 = str(tmp_path / "ours.gbrl_model")
    model_io.write_model(path2, meta, e, opts, name)
    m2 = ref.GBRL.load(path2)
    assert np.array_equal(np.asarray(m.predict(X, None)), np.asarray(m2.predict(X, None)))
    # byte-level: every section our writer emits equals the reference's file except the capacity fields of the metadata
    a, b = open(path, "rb").read(), open(path2, "rb").read()
    assert len(a) == len(b)
    diff = [i for i in range(len(a)) if a[i] != b[i]]
    # allowed: struct padding of the header (uninitialised in the reference) and max_trees .. max_leaves_batch
    assert all((20 <= i < 24) or (24 + 8 <= i < 24 + 24) or (6 <= i < 8) for i in diff), diff[:10]
    test_read_reference_file_and_write_it_back.keep = getattr(test_read_reference_file_and_write_it_back, "keep", []) + [ens, m, m2]
