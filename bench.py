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
c3 oblivious d8 cosine 4Mx64 D=2, c5 greedy d6 L2 8Mx256, c1 oblivious d4 10kx16.
"""
import argparse
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
    # PPO-minibatch shape driven through GBRL.step (shared actor-critic tree: 3 policy outputs + value), gbrl defaults
    "rl": dict(n=32_768, f=64, d=4, depth=4, grow="greedy", score="cosine", lrs=[(0.1, 0, 3), (0.01, 3, 4)]),
}
# predict-only workload (BASELINE config 4): 100k oblivious trees d6, D=2, batch 8192 x 128 (PPO rollout shape)
PREDICT = dict(n=8192, f=128, d=2, depth=6, n_trees=100_000, lrs=[(0.1, 0, 1), (0.01, 1, 2)])
# dram__bytes_read.sum + dram__bytes_write.sum of the histogram kernel, bytes per launch averaged over the six levels of a C2
# tree, from the committed `ncu --set full` capture (a number measured under a profiler cannot be taken live here)
NCU_TRAFFIC = {"c2": {"bytes_per_launch": 2.06e8,
                      "source": "profiles/r01_hist_full_c2.md: root level 269.2 MB, level 1 176.3 MB, level 2 197.6 MB "
                                "(levels 3-5 taken as level 2); algorithmic bytes of the same launches average 245 MB"}}
METRIC = "boosting-iters/sec (fit)"
UNIT = "iters/s"


def workload_name(w):
    c = WORKLOADS[w]
    return "%s: %s tree depth=%d %s score, %dx%d fp32, D=%d, quantile candidates, n_bins=256" % (
        w, c["grow"], c["depth"], c["score"], c["n"], c["f"], c["d"])


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


# ------------------------------------------------------------------------------------------------ CPU reference
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


def cpu_probe_rows(c, budget_s, steps):
    """Pick the sample size (rows) so that `steps` reference iterations take about budget_s seconds: the
    reference's cost is linear in N (O(d*B*F*N*D)), so probe a small N and scale."""
    n0 = 1024
    X, y = synth_numpy(n0, c["f"], c["d"], 0)
    kind, fit = cpu_reference_model(c, n0)
    t = time.perf_counter(); fit(X, y, 1); t = time.perf_counter() - t
    per_row = max(t, 1e-4) / n0
    rows = int(budget_s / max(steps, 1) / per_row)
    rows = max(512, min(rows, c["n"]))
    rows = 1 << int(math.log2(rows))
    return min(rows, c["n"]), kind


def run_cpu_sample(c, rows, warmup, steps):
    X, y = synth_numpy(rows, c["f"], c["d"], 0)
    kind, fit = cpu_reference_model(c, rows)
    if warmup > 0:
        fit(X, y, warmup)
    t = time.perf_counter(); fit(X, y, steps); t = time.perf_counter() - t
    its = steps / t
    return kind, its, its * rows / c["n"], t


# ------------------------------------------------------------------------------------------------ predict (config 4)
def bench_predict(args):
    """obs/s of the batched ensemble predict: 100k-tree oblivious ensemble, 8192 x 128 observations per call."""
    import numpy as np
    import torch
    from gbrl_b200 import GBRL
    p = PREDICT
    K, W = max(args.steps, 1), max(args.warmup, 0)
    rng = np.random.default_rng(0)
    nt, dep, f, d = p["n_trees"], p["depth"], p["f"], p["d"]
    nl = nt << dep
    e = {"tree_indices": (np.arange(nt, dtype=np.int64) << dep).astype(np.int32), "depths": np.full(nt, dep, np.int32),
         "values": (0.01 * rng.standard_normal((nl, d), dtype=np.float32)),
         "feature_indices": rng.integers(0, f, (nt, dep)).astype(np.int32),
         "feature_values": rng.standard_normal((nt, dep), dtype=np.float32),
         "edge_weights": np.zeros((nl, dep), np.float32), "inequality_directions": np.zeros((nl, dep), bool)}
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
    oh = np.empty((p["n"], d), np.float32)
    t = time.perf_counter()
    for _ in range(K):
        oh = m.predict_numpy(Xh.numpy())
    te = (time.perf_counter() - t) / K
    walks = p["n"] * nt
    out = {"metric": "obs/sec (predict)", "value": p["n"] / (ms * 1e-3), "unit": "obs/s", "n_gpus": 1, "steps": K, "warmup": W,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "c4: predict-only, %d oblivious trees depth %d, D=%d, batch %d obs x %d features" % (nt, dep, d, p["n"], f),
                      "l2": "ensemble arrays (%.0f MB) cycle through L2 every call" % ((nl * d * 4 + nt * dep * 8) / 1e6)},
           "tree_walks_per_s": walks / (ms * 1e-3),
           "e2e": {"value": p["n"] / te, "unit": "obs/s", "h2d_bytes_per_step": p["n"] * f * 4, "d2h_bytes_per_step": p["n"] * d * 4},
           "gpu_launches": int(launches)}
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------ step() API (RL path)
def bench_step_api(args):
    """GBRL_SB3's call pattern: every update calls predict(obs) then step(obs, grads) with device tensors; step()
    recomputes the quantile candidates from the batch (fitter.cpp:72-90), bins it and grows one tree."""
    import torch
    c = WORKLOADS["rl"]
    K, W = max(args.steps, 1), max(args.warmup, 0)
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
           "data": "synthetic", "config": {"workload": workload_name("rl") + ", device tensors, predict(all trees) + step per iteration"},
           "kernel_ms_per_step": {k: round(v["ms"] / K, 4) for k, v in prof.items() if isinstance(v, dict) and v["ms"] > 0},
           "gpu_launches": int(m.get_stats()["kernel_launches"] - l0), "replay": {k: m.get_stats()[k] for k in ("replay_nodes", "replay_items", "nodes_evaluated")}}
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c4"])   # "rl" is part of WORKLOADS
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-replay", action="store_true", help="exact-arithmetic arg-max only (see DESIGN.md, near-tie replay)")
    ap.add_argument("--hist-variant", type=int, default=0, help="0 streaming histogram kernel (default), 1 per-item kernel")
    ap.add_argument("--replay-variant", type=int, default=0, help="0 GPU-wide replay chains where output_dim <= 2 (default), 1 one CTA per replay item")
    ap.add_argument("--kappa", type=float, default=0.0, help="near-tie band width in noise units (0 = engine default)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds of CPU work for --impl reference")
    args = ap.parse_args()
    if args.workload == "c4":
        return bench_predict(args)
    if args.workload == "rl":
        return bench_step_api(args)
    c = WORKLOADS[args.workload]
    K, W = max(args.steps, 1), max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        rows, _ = cpu_probe_rows(c, args.ref_budget, K + W)
        kind, its_sample, its_full, t = run_cpu_sample(c, rows, W, K)
        sample = "%d of %d rows (full F=%d, depth, n_bins), %d timed boosting iterations in %.1f s; value scaled by rows/N (cost is linear in N)" % (
            rows, c["n"], c["f"], K, t)
        out = {"impl": "reference", "metric": METRIC, "value": its_full, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
               "ms_per_step": 1000.0 / its_full, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": {"workload": workload_name(args.workload)},
               "cpu_baseline": {"value": its_full, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
               "e2e": {"value": its_full, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(out))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the engine has no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl")
    X, y = synth_torch(c["n"], c["f"], c["d"], 0, dev)
    m = make_engine(c, local, ref_threads=cores, tie_replay=not args.no_replay, hist_variant=args.hist_variant, replay_variant=args.replay_variant, band_kappa=args.kappa)
    if world > 1:
        m.init_distributed()
    m.fit_begin(X, y, shuffle=False)              # bias, candidates, binning: once per fit (fitter.cpp:134-151)
    m.fit_iterate(W, sync=True)
    m.profile(True)
    l0 = m.get_stats()["kernel_launches"]
    rows0 = m.get_profile()["hist_rows"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record()
    m.fit_iterate(K, sync=False)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.time()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    prof = m.get_profile()
    launches = m.get_stats()["kernel_launches"] - l0
    m.profile(False)
    loss = m.fit_end()
    stats = m.get_stats()
    value = K / (ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (histogram): algorithmic bytes = rows scanned * (4F + 4D + 4), SURVEY 8d
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
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
    achieved = alg_bytes / (hist_ms * 1e-3) / 1e9 if hist_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC.get(args.workload, {}).get("bytes_per_launch") if (world == 1 and args.hist_variant == 0) else None,
                "traffic_source": NCU_TRAFFIC.get(args.workload, {}).get("source") if (world == 1 and args.hist_variant == 0) else None,
                "kernel": "hist_stream_kernel" if args.hist_variant == 0 else "hist_kernel", "launches": hist_launches,
                "avg_launch_ms": hist_ms / hist_launches,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": alg_bytes / hist_launches,
                "note": "algorithmic bytes count the fp32 row-major matrix (SURVEY 8d) for the rows the launch scans; the kernel "
                        "reads u16 codes (2 B/feature), so DRAM traffic is about half of that (see profiles/)"}
    breakdown = {k: round(v["ms"] / K, 4) for k, v in prof.items() if isinstance(v, dict) and v["ms"] > 0}

    # ---- e2e: the reference-facing call with HOST buffers (pinned), copies + candidates + binning inside
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((c["n"], c["f"]), dtype=torch.float32, pin_memory=True); Xh.copy_(X)
        yh = torch.empty((c["n"], c["d"]), dtype=torch.float32, pin_memory=True); yh.copy_(y)
        m2 = make_engine(c, local, ref_threads=cores, tie_replay=not args.no_replay, hist_variant=args.hist_variant, replay_variant=args.replay_variant, band_kappa=args.kappa)
        if world > 1:
            pass   # e2e is reported for rank 0's single-GPU call only when world > 1
        hx = (Xh.data_ptr(), tuple(Xh.shape), "torch.float32", "cpu")
        hy = (yh.data_ptr(), tuple(yh.shape), "torch.float32", "cpu")
        # untimed warm-up call (W boosting iterations): first-use allocation of the workspace, lazy module loading
        m2.fit(hx, None, hy, max(W, 1), False, "MultiRMSE")
        torch.cuda.synchronize()
        te = time.perf_counter()
        loss2 = m2.fit(hx, None, hy, K, False, "MultiRMSE")
        torch.cuda.synchronize()
        te = time.perf_counter() - te
        e2e = {"value": K / te, "unit": UNIT, "h2d_bytes_per_step": (Xh.numel() + yh.numel()) * 4 // K,
               "d2h_bytes_per_step": (4 + 4 * c["d"] + 256) // K + 1, "seconds": te, "loss": loss2,
               "call": "GBRL.fit(pinned host obs, pinned host targets, iterations=%d, shuffle=False) after one untimed warm-up call; "
                       "H2D copies, candidate generation, binning, bias and the final loss read-back are inside" % K}
        del m2, Xh, yh

    # ---- the same K iterations with the exact-arithmetic tier only (no reference-order replay of near-ties)
    exact_only = None
    if not args.no_replay and world == 1:
        m3 = make_engine(c, local, ref_threads=cores, tie_replay=False, hist_variant=args.hist_variant, replay_variant=args.replay_variant, band_kappa=args.kappa)
        m3.fit_begin(X, y, shuffle=False)
        m3.fit_iterate(W, sync=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m3.fit_iterate(K, sync=False); e1.record(); torch.cuda.synchronize()
        ms3 = e0.elapsed_time(e1)
        m3.fit_end()
        exact_only = {"value": K / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3 / K,
                      "note": "tie_replay=0: arg-max on exact integer-histogram sums only; differs from the reference only where "
                              "the reference's own sequential-fp32 rounding noise decides between near-tied candidates"}
        del m3

    # ---- CPU baseline on this box's host cores: bounded sample, scaled to the metric's unit
    cpu = None
    if not args.no_cpu_baseline and args.gpus == 1:
        try:
            rows, _ = cpu_probe_rows(c, args.cpu_budget, 2)
            kind, its_sample, its_full, t = run_cpu_sample(c, rows, 0, 2)
            cpu = {"value": its_full, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": "%d of %d rows (full F, depth, n_bins), 2 boosting iterations in %.1f s; scaled by rows/N" % (rows, c["n"], t)}
        except Exception as ex:   # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "unavailable", "sample": repr(ex)[:200]}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": ms / K,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 scores / int64 fixed-point sums",
           "data": "synthetic",
           "config": {"workload": workload_name(args.workload), "parallelism": "histogram sharded over %d rank(s): feature tiles x row chunks, one int64 all-reduce per level" % world,
                      "l2": "inputs larger than L2 (code matrix %.0f MB + fp32 matrix %.0f MB per level pass)" % (
                          c["n"] * c["f"] * 2 / 1e6, c["n"] * c["f"] * 4 / 1e6),
                      "tie_replay": not args.no_replay, "ref_threads": cores},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
           "kernel_ms_per_step": breakdown, "final_loss": loss, "exact_tier_only": exact_only,
           "replay": {"nodes": stats["replay_nodes"], "items": stats["replay_items"], "overflow": stats["replay_overflow"],
                      "nodes_evaluated": stats["nodes_evaluated"], "max_noise_ratio": stats["max_noise_ratio"],
                      "chain_blocks_fast": stats["chain_blocks_fast"], "chain_blocks_slow": stats["chain_blocks_slow"],
                      "chain_lanes_seq": stats["chain_lanes_seq"]}}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
