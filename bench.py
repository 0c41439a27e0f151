#!/usr/bin/env python
"""bench.py -- boosting-iterations/s of the fit hot path (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 3                       # this engine, BASELINE config[1] (C2)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1       # the reference's own CPU path (oracle/_ref)

A "step" is ONE boosting iteration of GBRL.fit (fitter.cpp:176-231: MultiRMSE gradients -> build_grads ->
grow one tree -> leaf values -> predictions) on a fixed synthetic matrix resident in HBM.  Candidate
generation and binning happen once per fit() call (fitter.cpp:134-151) and are outside `value`'s timed
region; they are inside `e2e`, which times the reference-facing call GBRL.fit(host arrays).

Workloads (BASELINE.json configs): c2 greedy d6 L2 1Mx128 (default, the config the metric is quoted on),
c3 oblivious d8 cosine 4Mx64 D=2, c5 greedy d6 L2 8Mx256, c1 oblivious d4 10kx16, j3 = north_star's "1M x 128
oblivious fit", c4 = predict-only (100k trees x 8192 obs), rl = PPO-minibatch shape through the step() API.
The default single-GPU run also measures c3 / c5 / j3 / c4 (each with its own roofline and bounded-sample CPU baseline)
and reports them under "extra_workloads"; a multi-GPU run adds c5 (the config the sharded histogram exists for).
"""
import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c1": dict(n=10_000, f=16, d=1, depth=4, grow="oblivious", score="L2", lrs=[(0.1, 0, 1)]),
    "c2": dict(n=1_000_000, f=128, d=1, depth=6, grow="greedy", score="L2", lrs=[(0.1, 0, 1)]),
    "c3": dict(n=4_000_000, f=64, d=2, depth=8, grow="oblivious", score="cosine", lrs=[(0.1, 0, 1), (0.01, 1, 2)]),
    "c5": dict(n=8_000_000, f=256, d=1, depth=6, grow="greedy", score="L2", lrs=[(0.1, 0, 1)]),
    # north_star's target workload: "1M x 128 oblivious fit"
    "j3": dict(n=1_000_000, f=128, d=1, depth=6, grow="oblivious", score="cosine", lrs=[(0.1, 0, 1)]),
    # PPO-minibatch shape driven through GBRL.step (shared actor-critic tree: 3 policy outputs + value), gbrl defaults
    "rl": dict(n=32_768, f=64, d=4, depth=4, grow="greedy", score="cosine", lrs=[(0.1, 0, 3), (0.01, 3, 4)]),
}
# predict-only workload (BASELINE config 4): 100k oblivious trees d6, D=2, batch 8192 x 128 (PPO rollout shape)
PREDICT = dict(n=8192, f=128, d=2, depth=6, n_trees=100_000, lrs=[(0.1, 0, 1), (0.01, 1, 2)])
METRIC = "boosting-iters/sec (fit)"
UNIT = "iters/s"
KEYS = ("tree_indices", "depths", "feature_indices", "feature_values", "inequality_directions", "edge_weights", "values")


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture of this round (a number measured under a profiler cannot be taken live inside a timed run)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return t.get(workload)
    except Exception:
        return None


def workload_name(w):
    c = WORKLOADS[w]
    return "%s: %s tree depth=%d %s score, %dx%d fp32, D=%d, quantile candidates, n_bins=256" % (
        w, c["grow"], c["depth"], c["score"], c["n"], c["f"], c["d"])


def predict_name():
    p = PREDICT
    return "c4: predict-only, %d oblivious trees depth %d, D=%d, batch %d obs x %d features" % (p["n_trees"], p["depth"], p["d"], p["n"], p["f"])


def peak_hbm():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.p, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], None, set()
        for (ts, line) in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            parts = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(parts[1])); mx = float(parts[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:      # region shorter than one sample: fall back to every sample we have
            for (ts, line) in self.rows:
                parts = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(parts[1])); mx = float(parts[2])
                except Exception:
                    pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ data
def synth_torch(n, f, d, seed, device):
    """X ~ N(0,1), targets = tanh(XW/sqrt(F)) + 0.1 eps; generated on the device, identical on every rank."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    X = torch.randn((n, f), generator=g, device=device, dtype=torch.float32)
    W = torch.randn((f, d), generator=g, device=device, dtype=torch.float32)
    y = torch.tanh(X @ W / math.sqrt(f)) + 0.1 * torch.randn((n, d), generator=g, device=device, dtype=torch.float32)
    return X.contiguous(), y.contiguous()


def synth_numpy(n, f, d, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, f), dtype=np.float32)
    W = rng.standard_normal((f, d), dtype=np.float32)
    y = (np.tanh(X @ W / np.sqrt(f)) + 0.1 * rng.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
    return X, y


def make_engine(c, device_index, ref_threads, tie_replay=True, hist_variant=0, replay_variant=0, band_kappa=0.0):
    import numpy as np
    from gbrl_b200 import GBRL
    m = GBRL(input_dim=c["f"], output_dim=c["d"], policy_dim=c["d"], max_depth=c["depth"], n_bins=256, par_th=10,
             split_score_func=c["score"], generator_type="quantile", batch_size=c["n"], grow_policy=c["grow"],
             device="cuda:%d" % device_index, ref_threads=ref_threads, tie_replay=tie_replay, hist_variant=hist_variant, replay_variant=replay_variant, band_kappa=band_kappa)
    m.set_bias(np.zeros(c["d"], np.float32))
    m.set_feature_weights(np.ones(c["f"], np.float32))
    m.set_feature_mapping(np.arange(c["f"], dtype=np.int32), np.ones(c["f"], dtype=bool))
    for (lr, a, b) in c["lrs"]:
        m.set_optimizer("SGD", "const", lr, a, b)
    return m


def predict_ensemble(rng):
    """The synthetic C4 ensemble: random feature in [0,128), threshold ~ N(0,1), values ~ N(0, 0.01)."""
    import numpy as np
    p = PREDICT
    nt, dep, f, d = p["n_trees"], p["depth"], p["f"], p["d"]
    nl = nt << dep
    li = np.arange(1 << dep)
    iq = ((li[:, None] >> (dep - 1 - np.arange(dep))[None, :]) & 1).astype(bool)
    return {"tree_indices": (np.arange(nt, dtype=np.int64) << dep).astype(np.int32), "depths": np.full(nt, dep, np.int32),
            "values": (0.01 * rng.standard_normal((nl, d), dtype=np.float32)),
            "feature_indices": rng.integers(0, f, (nt, dep)).astype(np.int32),
            "feature_values": rng.standard_normal((nt, dep), dtype=np.float32),
            "edge_weights": np.zeros((nl, dep), np.float32), "inequality_directions": np.tile(iq, (nt, 1))}


# ------------------------------------------------------------------------------------------------ CPU reference (child process)
def cpu_reference_model(c, n_rows):
    """The reference's own CPU path (oracle/_ref, compiled from the reference sources) or, if that is not
    present, the plain-C oracle port.  Returns (kind, fit_callable)."""
    import numpy as np
    from oracle.oracle import Oracle, load_reference, make_reference
    ref = load_reference()
    kw = dict(input_dim=c["f"], output_dim=c["d"], max_depth=c["depth"], n_bins=256, par_th=10, split_score_func=c["score"],
              generator_type="quantile", batch_size=n_rows, grow_policy=c["grow"])
    if ref is not None:
        m = make_reference(ref, lrs=c["lrs"], **kw)
        return "reference", (lambda X, y, it: m.fit(X, None, y, it, False, "MultiRMSE"))
    o = Oracle(ref_threads=os.cpu_count() or 1, **kw)
    o.set_bias(np.zeros(c["d"], np.float32)); o.set_feature_weights(np.ones(c["f"], np.float32))
    o.set_feature_mapping(np.arange(c["f"]), np.ones(c["f"]))
    for (lr, a, b) in c["lrs"]:
        o.set_optimizer("SGD", "const", lr, a, b)
    return "port", (lambda X, y, it: o.fit(X, y, it))


def cpu_child(args):
    """Runs in a CHILD process whose OMP_NUM_THREADS was set explicitly by the parent (torchrun exports OMP_NUM_THREADS=1
    to its workers, and the reference's thread count is latched when libgomp loads): probes the cost per row, picks the
    sample size for the budget (the reference's cost is linear in N), times `steps` boosting iterations (fit
    workloads) or predict calls (c4) and prints one JSON object."""
    import numpy as np
    K, W = max(args.steps, 1), max(args.warmup, 0)
    if args.workload == "c4":
        from oracle.oracle import load_reference
        from gbrl_b200 import model_io
        ref = load_reference()
        p = PREDICT
        rng = np.random.default_rng(0)
        e = predict_ensemble(rng)
        nt, dep, f, d = p["n_trees"], p["depth"], p["f"], p["d"]
        e.update({"bias": np.zeros(d, np.float32), "feature_weights": np.ones(f, np.float32),
                  "reverse_num_feature_mapping": np.arange(f, dtype=np.int32), "reverse_cat_feature_mapping": np.full(f, -1, np.int32),
                  "feature_mapping": np.arange(f, dtype=np.int32), "mapping_numerics": np.ones(f, bool)})
        meta = {"n_leaves": nt << dep, "n_trees": nt, "input_dim": f, "output_dim": d, "policy_dim": d, "max_depth": dep,
                "min_data_in_leaf": 0, "n_bins": 256, "par_th": 10, "cv_beta": 0.9, "verbose": 0, "batch_size": p["n"], "use_cv": 0,
                "split_score_func": 1, "generator_type": 1, "grow_policy": 1, "n_num_features": f, "n_cat_features": 0, "iteration": nt}
        opts = [{"algo": "SGD", "scheduler_func": "Const", "init_lr": lr, "start_idx": a, "stop_idx": b, "stop_lr": 1e-8, "T": 10000}
                for (lr, a, b) in p["lrs"]]
        path = "/tmp/bench_c4_%d.gbrl_model" % os.getpid()
        model_io.write_model(path, meta, e, opts, "GBRL")
        m = ref.GBRL.load(path)
        os.unlink(path)
        rows = args.rows if args.rows > 0 else 1024
        X = rng.standard_normal((rows, f), dtype=np.float32)
        m.predict(X[:64], None)
        t = time.perf_counter()
        for _ in range(K):
            m.predict(X, None)
        t = time.perf_counter() - t
        print(json.dumps({"kind": "reference", "rows": rows, "seconds": t, "steps": K, "obs_per_s": rows * K / t}), flush=True)
        os._exit(0)
    c = WORKLOADS[args.workload]
    rows = args.rows
    if rows <= 0:
        n0 = 1024
        X, y = synth_numpy(n0, c["f"], c["d"], 0)
        kind, fit = cpu_reference_model(c, n0)
        t = time.perf_counter(); fit(X, y, 1); t = time.perf_counter() - t
        per_row = max(t, 1e-4) / n0
        rows = int(args.budget / max(K + W, 1) / per_row)
        rows = max(512, min(rows, c["n"]))
        rows = min(1 << int(math.log2(rows)), c["n"])
    X, y = synth_numpy(rows, c["f"], c["d"], 0)
    kind, fit = cpu_reference_model(c, rows)
    if W > 0:
        fit(X, y, W)
    t = time.perf_counter(); fit(X, y, K); t = time.perf_counter() - t
    print(json.dumps({"kind": kind, "rows": rows, "seconds": t, "steps": K, "its_sample": K / t}), flush=True)
    os._exit(0)      # the reference module's teardown is not worth waiting for


def run_cpu_sample(workload, budget_s, steps, warmup, rows=0):
    """Spawns cpu_child with every host core.  Returns dict(kind, rows, seconds, value [metric unit at full size], cores)."""
    cores = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    for k in ("OMP_PROC_BIND", "OMP_PLACES", "GOMP_CPU_AFFINITY", "KMP_AFFINITY"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-child", "--workload", workload, "--budget", str(budget_s), "--steps", str(steps),
           "--warmup", str(warmup), "--rows", str(rows)]
    try:
        os.sched_setaffinity(0, range(cores))      # torchrun may have pinned this rank; the child inherits the mask
    except Exception:
        pass
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=3600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        raise RuntimeError("cpu sample failed: " + (r.stderr or r.stdout)[-400:])
    d = json.loads(lines[-1])
    d["cores"] = cores
    if workload == "c4":
        d["value"] = d["obs_per_s"]
        d["sample"] = "%d of %d observations x all %d trees, %d predict calls in %.1f s (reference predict is linear in observations)" % (
            d["rows"], PREDICT["n"], PREDICT["n_trees"], d["steps"], d["seconds"])
    else:
        n = WORKLOADS[workload]["n"]
        d["value"] = d["its_sample"] * d["rows"] / n
        full = d["rows"] >= n
        d["sample"] = ("%d of %d rows (full F=%d, depth, n_bins), %d timed boosting iterations in %.1f s%s" % (
            d["rows"], n, WORKLOADS[workload]["f"], d["steps"], d["seconds"],
            "" if full else "; value scaled by rows/N (cost is linear in N)"))
    return d


# ------------------------------------------------------------------------------------------------ predict (config 4)
def bench_predict(args, K, W, with_cpu=True):
    """obs/s of the batched ensemble predict: 100k-tree oblivious ensemble, 8192 x 128 observations per call."""
    import numpy as np
    import torch
    from gbrl_b200 import GBRL
    p = PREDICT
    rng = np.random.default_rng(0)
    nt, dep, f, d = p["n_trees"], p["depth"], p["f"], p["d"]
    nl = nt << dep
    e = predict_ensemble(rng)
    m = GBRL(input_dim=f, output_dim=d, policy_dim=d, max_depth=dep, n_bins=256, split_score_func="cosine",
             generator_type="quantile", batch_size=p["n"], grow_policy="oblivious", device="cuda:0")
    m.set_bias(np.zeros(d, np.float32)); m.set_feature_weights(np.ones(f, np.float32))
    m.set_feature_mapping(np.arange(f, dtype=np.int32), np.ones(f, dtype=bool))
    for (lr, a, b) in p["lrs"]:
        m.set_optimizer("SGD", "const", lr, a, b)
    m._set_ensemble(e, f)
    X = torch.randn((p["n"], f), device="cuda", dtype=torch.float32)
    for _ in range(W):
        m.predict_tensor(X)
    torch.cuda.synchronize()
    l0 = m.get_stats()["kernel_launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        out = m.predict_tensor(X)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / K
    launches = m.get_stats()["kernel_launches"] - l0
    Xh = torch.empty((p["n"], f), dtype=torch.float32, pin_memory=True); Xh.copy_(X)
    t = time.perf_counter()
    for _ in range(K):
        m.predict_numpy(Xh.numpy())
    te = (time.perf_counter() - t) / K
    walks = p["n"] * nt
    peak, peak_src = peak_hbm()
    ens_bytes = nl * d * 4 + nt * dep * 8
    alg = ens_bytes + p["n"] * f * 4 + p["n"] * d * 4
    # honest bound: the work unit is the tree walk -- depth x (shared-memory feature read + compare) + one value
    # gather of D floats + D multiply-subtracts, ~ (4 * depth + 2 * D + 6) issue slots per 32 walks; the ensemble is
    # L2-resident, so the HBM figure only says how little of the time is DRAM
    slots = 4 * dep + 2 * d + 6
    issue_peak = 148 * 4 * 1.965e9 / slots * 32
    out = {"metric": "obs/sec (predict)", "value": p["n"] / (ms * 1e-3), "unit": "obs/s", "n_gpus": 1, "steps": K, "warmup": W,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": predict_name(), "l2": "ensemble arrays (%.0f MB) cycle through L2 every call" % (ens_bytes / 1e6)},
           "tree_walks_per_s": walks / (ms * 1e-3),
           "e2e": {"value": p["n"] / te, "unit": "obs/s", "h2d_bytes_per_step": p["n"] * f * 4, "d2h_bytes_per_step": p["n"] * d * 4},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                        "traffic": (ncu_traffic("c4") or {}).get("bytes_per_launch"), "peak_source": peak_src, "kernel": "predict_tiles_kernel",
                        "algorithmic_bytes_per_launch": alg,
                        "note": "predict is not HBM-bound (ensemble + observations sit in L2, SURVEY 8d); the bound that matters is the "
                                "issue rate of the tree walk, reported under issue_bound"},
           "issue_bound": {"walks_per_s": walks / (ms * 1e-3), "peak_walks_per_s": issue_peak, "frac": walks / (ms * 1e-3) / issue_peak,
                           "model": "%d issue slots per warp-walk (32 observations x 1 tree), 148 SMs x 4 schedulers x 1.965 GHz" % slots}}
    if with_cpu:
        try:
            cpu = run_cpu_sample("c4", 0, 2, 0, rows=1024)
            out["cpu_baseline"] = {"value": cpu["value"], "unit": "obs/s", "cores": cpu["cores"], "kind": cpu["kind"], "sample": cpu["sample"]}
        except Exception as ex:   # pragma: no cover
            out["cpu_baseline"] = {"value": None, "unit": "obs/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(ex)[:200]}
    del m, X
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ step() API (RL path)
def bench_step_api(args, K, W, workload="rl"):
    """GBRL_SB3's call pattern: every update calls predict(obs) then step(obs, grads) with device tensors; step()
    recomputes the quantile candidates from the batch (fitter.cpp:72-90), bins it and grows one tree."""
    import torch
    c = WORKLOADS[workload]
    dev = torch.device("cuda", 0)
    X, y = synth_torch(c["n"], c["f"], c["d"], 0, dev)
    m = make_engine(c, 0, ref_threads=os.cpu_count() or 1, tie_replay=not args.no_replay)
    ti = lambda t: (t.data_ptr(), tuple(t.shape), "torch.float32", "cuda")

    def one():
        p = torch.from_dlpack(m.predict(ti(X), None))
        g = (p.reshape(c["n"], c["d"]) - y).contiguous()
        m.step(ti(X), None, ti(g))
    for _ in range(W):
        one()
    m.profile(True)
    torch.cuda.synchronize()
    l0 = m.get_stats()["kernel_launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        one()
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / K
    prof = m.get_profile()
    out = {"metric": "boosting-iters/sec (step API: predict + step per call)", "value": 1000.0 / ms, "unit": UNIT, "n_gpus": 1,
           "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": {"workload": workload_name(workload) + ", device tensors, predict(all trees) + step per iteration"},
           "kernel_ms_per_step": {k: round(v["ms"] / K, 4) for k, v in prof.items() if isinstance(v, dict) and v["ms"] > 0},
           "gpu_launches": int(m.get_stats()["kernel_launches"] - l0), "replay": {k: m.get_stats()[k] for k in ("replay_nodes", "replay_items", "nodes_evaluated")}}
    return out


# ------------------------------------------------------------------------------------------------ fit workloads
def ensemble_sha(m):
    import numpy as np
    e = m.get_ensemble_data()
    h = hashlib.sha256()
    for k in KEYS:
        h.update(np.ascontiguousarray(e[k]).tobytes())
    return h.hexdigest()[:16]


def bench_fit(wl, args, K, W, rank, world, local, with_e2e=True, with_exact=True, with_cpu=True, cpu_budget=20.0):
    """One fit workload on `world` ranks.  Returns the JSON dict on rank 0, None elsewhere."""
    import numpy as np
    import torch
    import torch.distributed as dist
    c = WORKLOADS[wl]
    cores = os.cpu_count() or 1
    dev = torch.device("cuda", local)
    X, y = synth_torch(c["n"], c["f"], c["d"], 0, dev)
    m = make_engine(c, local, ref_threads=cores, tie_replay=not args.no_replay, hist_variant=args.hist_variant, replay_variant=args.replay_variant, band_kappa=args.kappa)
    if world > 1:
        m.init_distributed()
    m.fit_begin(X, y, shuffle=False)              # bias, candidates, binning: once per fit (fitter.cpp:134-151)
    m.fit_iterate(W, sync=True)
    l0 = m.get_stats()["kernel_launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.time()
    ev0.record()
    m.fit_iterate(K, sync=False)                  # the timed region of `value`: exactly K boosting iterations
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = m.get_stats()["kernel_launches"] - l0
    # the next K iterations of the same fit, with the engine's per-kernel-class CUDA events on (two events around every
    # class of launches cost ~0.2 ms per iteration, measured, so the headline region above runs without them): kernel
    # durations, roofline and the per-class breakdown come from this second region
    m.profile(True)
    rows0 = m.get_profile()["hist_rows"]
    m.fit_iterate(K, sync=False)
    ev2.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.time()
    ms = ev0.elapsed_time(ev1)
    ms_prof = ev1.elapsed_time(ev2)
    if world > 1:
        t = torch.tensor([ms, ms_prof], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_prof = float(t[0].item()), float(t[1].item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    prof = m.get_profile()
    m.profile(False)
    loss = m.fit_end()
    stats = m.get_stats()
    sha = ensemble_sha(m)
    if world > 1:      # every rank must hold the same ensemble (integer histograms: bit-identical by construction)
        shas = [None] * world
        dist.all_gather_object(shas, sha)
        assert all(s == shas[0] for s in shas), "ranks disagree on the ensemble: %s" % shas
    value = K / (ms * 1e-3)
    del m

    # ---- e2e: the reference-facing call with HOST buffers (pinned), copies + candidates + binning inside; on N ranks the
    #      same call on every rank (each holds all rows; histograms sharded), timed between barriers, max over ranks
    e2e = None
    if with_e2e:
        Xh = torch.empty((c["n"], c["f"]), dtype=torch.float32, pin_memory=True); Xh.copy_(X)
        yh = torch.empty((c["n"], c["d"]), dtype=torch.float32, pin_memory=True); yh.copy_(y)
        m2 = make_engine(c, local, ref_threads=cores, tie_replay=not args.no_replay, hist_variant=args.hist_variant, replay_variant=args.replay_variant, band_kappa=args.kappa)
        if world > 1:
            m2.init_distributed()
        hx = (Xh.data_ptr(), tuple(Xh.shape), "torch.float32", "cpu")
        hy = (yh.data_ptr(), tuple(yh.shape), "torch.float32", "cpu")
        # untimed warm-up call (W boosting iterations): first-use allocation of the workspace, lazy module loading
        m2.fit(hx, None, hy, max(W, 1), False, "MultiRMSE")
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        te = time.perf_counter()
        loss2 = m2.fit(hx, None, hy, K, False, "MultiRMSE")
        torch.cuda.synchronize()
        te = time.perf_counter() - te
        if world > 1:
            t = torch.tensor([te], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        e2e = {"value": K / te, "unit": UNIT, "h2d_bytes_per_step": (Xh.numel() + yh.numel()) * 4 // K,
               "d2h_bytes_per_step": (4 + 4 * c["d"] + 256) // K + 1, "seconds": te, "loss": loss2,
               "call": "GBRL.fit(pinned host obs, pinned host targets, iterations=%d, shuffle=False) after one untimed warm-up call%s; "
                       "H2D copies, candidate generation, binning, bias and the final loss read-back are inside" % (
                           K, "" if world == 1 else ", on every one of the %d ranks (max over ranks)" % world)}
        del m2, Xh, yh

    # ---- the same K iterations with the exact-arithmetic tier only (no reference-order replay of near-ties)
    exact_only = None
    if with_exact and not args.no_replay and world == 1:
        m3 = make_engine(c, local, ref_threads=cores, tie_replay=False, hist_variant=args.hist_variant, replay_variant=args.replay_variant, band_kappa=args.kappa)
        m3.fit_begin(X, y, shuffle=False)
        m3.fit_iterate(W, sync=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m3.fit_iterate(K, sync=False); e1.record(); torch.cuda.synchronize()
        ms3 = e0.elapsed_time(e1)
        m3.profile(True)
        rows3 = m3.get_profile()["hist_rows"]
        m3.fit_iterate(min(K, 10), sync=True)
        prof3 = m3.get_profile()
        m3.profile(False)
        m3.fit_end()
        exact_only = {"value": K / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3 / K,
                      "note": "tie_replay=0: arg-max on exact integer-histogram sums only; differs from the reference only where "
                              "the reference's own sequential-fp32 rounding noise decides between near-tied candidates",
                      "hist_ms": prof3["histogram"]["ms"], "hist_launches": prof3["histogram"]["launches"], "hist_rows": prof3["hist_rows"] - rows3}
        del m3
    del X, y
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel (histogram): algorithmic bytes = rows scanned * (4F + 4D + 4), SURVEY 8d
    peak, peak_src = peak_hbm()
    hist_ms = prof["histogram"]["ms"]; hist_launches = max(prof["histogram"]["launches"], 1)
    rows_scanned = prof["hist_rows"] - rows0
    n_tiles = (c["f"] + 31) // 32
    gt = min(world, n_tiles)                      # same 2-D sharding arithmetic as prepare_workspace() in capi.cu
    while gt > 1 and world % gt != 0:
        gt -= 1
    tg = rank % gt
    own_tiles = (n_tiles * (tg + 1)) // gt - (n_tiles * tg) // gt      # feature tiles this rank histograms
    f_local = min(c["f"], own_tiles * 32)
    alg_bytes = rows_scanned * (4 * f_local + 4 * c["d"] + 4)
    dram_bytes = rows_scanned * (2 * f_local + 4 * c["d"] + 4)         # what the kernel really streams: u16 codes, gradient, row id
    achieved = alg_bytes / (hist_ms * 1e-3) / 1e9 if hist_ms > 0 else 0.0
    tr = ncu_traffic(wl) if (world == 1 and args.hist_variant == 0) else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_dram": (dram_bytes / (hist_ms * 1e-3) / 1e9 / peak) if hist_ms > 0 else 0.0,
                "traffic": (tr or {}).get("bytes_per_launch"), "traffic_source": (tr or {}).get("source"),
                "kernel": "hist_stream_kernel" if args.hist_variant == 0 else "hist_kernel", "launches": hist_launches,
                "avg_launch_ms": hist_ms / hist_launches, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / hist_launches,
                "note": "frac counts the fp32 row-major matrix (SURVEY 8d: rows scanned x (4F + 4D + 4)); frac_dram counts the bytes the "
                        "kernel really streams (u16 codes: rows x (2F + 4D + 4))"}
    if exact_only and exact_only.get("hist_ms", 0) > 0 and world == 1:
        # the same kernel with nothing else on the GPU: the speculative replay of the main run shares the SMs with it
        b3 = exact_only.pop("hist_rows") * (4 * f_local + 4 * c["d"] + 4)
        h3 = exact_only.pop("hist_ms"); n3 = max(exact_only.pop("hist_launches"), 1)
        roofline["no_overlap"] = {"achieved": b3 / (h3 * 1e-3) / 1e9, "frac": b3 / (h3 * 1e-3) / 1e9 / peak, "avg_launch_ms": h3 / n3,
                                  "note": "histogram launches of the exact-tier-only run (no replay kernels on side streams)"}
    breakdown = {k: round(v["ms"] / K, 4) for k, v in prof.items() if isinstance(v, dict) and v["ms"] > 0}

    # ---- CPU baseline on this box's host cores: bounded sample, scaled to the metric's unit
    cpu = None
    if with_cpu:
        try:
            s = run_cpu_sample(wl, cpu_budget, 2, 0)
            cpu = {"value": s["value"], "unit": UNIT, "cores": s["cores"], "kind": s["kind"], "sample": s["sample"]}
        except Exception as ex:   # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "unavailable", "sample": repr(ex)[:200]}

    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 scores / int64 fixed-point sums",
            "data": "synthetic",
            "config": {"workload": workload_name(wl), "parallelism": "histogram sharded over %d rank(s): feature tiles x row chunks, one int64 all-reduce of the level buffer per level" % world,
                       "l2": "inputs larger than L2 (code matrix %.0f MB + fp32 matrix %.0f MB per level pass)" % (
                           c["n"] * c["f"] * 2 / 1e6, c["n"] * c["f"] * 4 / 1e6),
                       "tie_replay": not args.no_replay, "ref_threads": cores},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "kernel_ms_per_step": breakdown, "profiled_pass": {"ms_per_step": ms_prof / K, "steps": K,
                "note": "iterations K..2K of the same fit with per-kernel-class CUDA events enabled: source of kernel_ms_per_step and roofline"},
            "final_loss": loss, "ensemble_sha": sha, "exact_tier_only": exact_only,
            "replay": {"nodes": stats["replay_nodes"], "items": stats["replay_items"], "overflow": stats["replay_overflow"],
                       "nodes_evaluated": stats["nodes_evaluated"], "max_noise_ratio": stats["max_noise_ratio"],
                       "chain_blocks_fast": stats["chain_blocks_fast"], "chain_blocks_slow": stats["chain_blocks_slow"],
                       "chain_lanes_seq": stats["chain_lanes_seq"], "flips": stats["replay_flips"],
                       "speculative_trees": stats["spec_trees"], "levels_rolled_back": stats["spec_rollbacks"]}}


def compact(d):
    """The part of a workload's line that is kept under extra_workloads."""
    keep = ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches",
            "kernel_ms_per_step", "profiled_pass", "ensemble_sha", "tree_walks_per_s", "issue_bound", "exact_tier_only", "replay")
    return {k: d[k] for k in keep if k in d and d[k] is not None}


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c4"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload (no extra_workloads)")
    ap.add_argument("--no-replay", action="store_true", help="exact-arithmetic arg-max only (see DESIGN.md, near-tie replay)")
    ap.add_argument("--hist-variant", type=int, default=0, help="0 streaming histogram kernel (default), 1 per-item kernel")
    ap.add_argument("--replay-variant", type=int, default=0, help="bit 0: 0 GPU-wide replay chains where output_dim <= 2 (default), 1 one CTA per replay item; bit 1: 0 speculative levels (default), 1 every level waits for its replay")
    ap.add_argument("--kappa", type=float, default=0.0, help="near-tie band width in noise units (0 = engine default)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds of CPU work for --impl reference")
    ap.add_argument("--ref-rows", type=int, default=0, help="--impl reference: force the sample size (e.g. the full N for a directly timed iteration)")
    ap.add_argument("--extras-budget", type=float, default=170.0, help="wall-clock seconds after which remaining extra workloads are skipped")
    # internal: the CPU sample runs in a child process with an explicit OMP_NUM_THREADS (see cpu_child)
    ap.add_argument("--cpu-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--budget", type=float, default=20.0, help=argparse.SUPPRESS)
    ap.add_argument("--rows", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_child:
        return cpu_child(args)
    K, W = max(args.steps, 1), max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    t_start = time.time()

    if args.impl == "reference":
        if rank != 0:
            return 0
        wl = args.workload
        s = run_cpu_sample(wl, args.ref_budget, K, W, rows=args.ref_rows)
        unit = "obs/s" if wl == "c4" else UNIT
        out = {"impl": "reference", "metric": "obs/sec (predict)" if wl == "c4" else METRIC, "value": s["value"], "unit": unit, "n_gpus": args.gpus,
               "steps": K, "warmup": W, "ms_per_step": 1000.0 / s["value"] if wl != "c4" else 1000.0 * PREDICT["n"] / s["value"],
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": {"workload": predict_name() if wl == "c4" else workload_name(wl)},
               "cpu_baseline": {"value": s["value"], "unit": unit, "cores": s["cores"], "kind": s["kind"], "sample": s["sample"]},
               "e2e": {"value": s["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(out))
        return 0

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the engine has no CPU fallback)"
    torch.cuda.set_device(local)
    if args.workload == "c4":
        print(json.dumps(bench_predict(args, K, W, with_cpu=not args.no_cpu_baseline)))
        return 0
    if args.workload == "rl":
        print(json.dumps(bench_step_api(args, K, W)))
        return 0
    if world > 1:
        dist.init_process_group("nccl")
    out = bench_fit(args.workload, args, K, W, rank, world, local, with_e2e=not args.no_e2e, with_exact=True,
                    with_cpu=(not args.no_cpu_baseline) and args.gpus == 1, cpu_budget=args.cpu_budget)
    if rank == 0:
        out["n_gpus"] = args.gpus

    # ---- the other BASELINE configs (driver-visible: extra keys of the same line)
    extras = {}
    if not args.no_extras and args.workload == "c2":
        if world > 1:
            # the config the sharded histogram exists for (BASELINE config 5): every rank takes part
            Kx = min(K, 10)
            try:
                r = bench_fit("c5", args, Kx, min(W, 3), rank, world, local, with_e2e=False, with_exact=False, with_cpu=False)
                if rank == 0:
                    extras["c5"] = compact(r)
            except Exception as ex:   # pragma: no cover
                if rank == 0:
                    extras["c5"] = {"error": repr(ex)[:300]}
        else:
            for wl in ("j3", "c3", "c4", "c5"):
                if time.time() - t_start > args.extras_budget:
                    extras[wl] = {"skipped": "extras budget of %.0f s spent" % args.extras_budget}
                    continue
                try:
                    if wl == "c4":
                        r = bench_predict(args, min(K, 20), min(W, 3), with_cpu=not args.no_cpu_baseline)
                    else:
                        r = bench_fit(wl, args, min(K, 10), min(W, 3), 0, 1, local, with_e2e=(wl == "j3") and not args.no_e2e, with_exact=False,
                                      with_cpu=not args.no_cpu_baseline, cpu_budget=8.0)
                    extras[wl] = compact(r)
                except Exception as ex:   # pragma: no cover
                    extras[wl] = {"error": repr(ex)[:300]}
    if rank == 0:
        if extras:
            out["extra_workloads"] = extras
        out["bench_seconds"] = round(time.time() - t_start, 1)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
