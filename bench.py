#!/usr/bin/env python
"""bench.py -- candidate-fits/sec of the cross-validated grid search (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

A "step" is one pass of the hot path over the whole workload: every (candidate, fold) fit+score task of
BASELINE config 2 -- GridSearchCV(SVC rbf) on synthetic 10000x512 fp32, C x gamma 8x8, cv=5 = 320 fits per
GPU.  At N > 1 GPUs every rank holds the dataset and evaluates its own 64 candidates of a grid refined
on the same ranges (N*64 candidates dealt to the ranks by predicted cost): weak scaling, no data-path collective, one
all-gather of the score blocks per step (the counterpart of RDD.collect()).

value  : fits/s with the dataset resident in HBM (gs_set_data done before the timed region); device time
         from CUDA events recorded on the engine's stream around each gs_svc call, max over ranks.
e2e    : the same metric through the public API -- GridSearchCV(...).fit(X, y) with HOST numpy buffers
         every step (H2D of X/y/folds and D2H of the score arrays inside the timed region), refit=False.
roofline: the dominant kernel (batched SMO): algorithmic HBM bytes (2 gathered float32 K rows of the
         sub-problem per SMO iteration) / CUDA-event duration of the solve phase.
cpu_baseline / --impl reference: scikit-learn's own GridSearchCV (the reference's CPU path, joblib with
         all host cores) on a bounded, strided sample of the same candidate list.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------- helpers ----------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def scaled_workload(key, n_gpus):
    """Weak scaling: 64 candidates per GPU on the ranges of config 2 (N=1 is config 2 itself; N=4 is the
    16x16 grid of config 4)."""
    from spark_sklearn_b200 import workloads as W
    w = W.make_workload(key)
    if key == "c2" and n_gpus > 1:
        nc, ng = {2: (8, 16), 4: (16, 16), 8: (16, 32)}.get(n_gpus, (8, 8 * n_gpus))
        w["param_grid"] = {"C": np.logspace(-1, 2.5, nc), "gamma": np.geomspace(1 / 4096, 1 / 256, ng)}
        w["name"] = "c2_weak_%dx%d" % (nc, ng)
    return w


def effective_cores():
    """Host cores this process may actually use: min(affinity, cgroup cpu.max quota)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(round(int(quota) / int(period)))))
    except Exception:
        pass
    return n


def cpu_sample(cands, n_splits, steps_total):
    """Bounded sample of the candidate list for the CPU arm.  One libsvm fit of config 2 takes 25-50 s on the GPU
    box's 16-CPU cgroup, so the sample is two candidates from the middle of the (C, gamma) grid whose mean SMO
    iteration count (~18k) matches the grid mean (~16k): 10 fits, one wave on 16 cores."""
    n = len(cands)
    side = int(round(n ** 0.5))
    if side * side == n and side >= 4:
        idx = [(3 * side // 8) * side + (2 * side // 8), (6 * side // 8) * side + (3 * side // 8)]
    else:
        idx = [n // 3, (2 * n) // 3]
    if steps_total > 4:
        idx = idx[:1]                                  # many steps asked: one mid-grid candidate (5 fits, ~25 s) per step
    return idx, "%d of %d candidates (mid-grid, mean cost ~ grid mean), all %d folds: %d fits" % (
        len(idx), n, n_splits, len(idx) * n_splits)


def run_reference_step(w, cand_idx, cores):
    """The reference's CPU implementation of the path: sklearn GridSearchCV -> joblib -> _fit_and_score
    (what spark_sklearn maps over Spark executors, base_search.py:74-90), all host cores, refit=False."""
    from sklearn.model_selection import GridSearchCV
    from spark_sklearn_b200 import workloads as W
    cands = W.candidates(w)
    grid = [{k: [v] for k, v in cands[i].items()} for i in cand_idx]
    s = GridSearchCV(W.make_estimator(w), grid, cv=w["cv"], return_train_score=True, refit=False, n_jobs=cores)
    t0 = time.perf_counter()
    s.fit(w["X"], w["y"])
    dt = time.perf_counter() - t0
    return dt, len(cand_idx) * s.n_splits_, s.cv_results_["mean_test_score"]


# ----------------------------------------------------------------------------- main -------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner, joblib) go to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = effective_cores()
    W_ = max(a.warmup, 0)

    w = scaled_workload(a.workload, max(a.gpus, world))
    from spark_sklearn_b200 import workloads as WL
    cands = WL.candidates(w)
    n_splits = w["cv"]
    cfg = {"workload": "%s: GridSearchCV(%s %s), synthetic %dx%d fp32, %d candidates x cv=%d"
                       % (w["name"], w["estimator"], w["est_params"], w["X"].shape[0], w["X"].shape[1], len(cands), n_splits),
           "n_candidates": len(cands), "n_splits": n_splits, "fits_per_step": len(cands) * n_splits,
           "parallelism": "candidates dealt by predicted cost over %d GPU(s), dataset replicated, one score all-gather" % max(world, 1),
           "l2": "inputs larger than L2 (float64 Gram 0.8 GB + float32 kernel matrices 0.4 GB each)", "refit": False}

    # ---------------- reference arm: the CPU path on the host cores (rank 0 only) ----------------
    if a.impl == "reference":
        if rank != 0:
            return
        idx, desc = cpu_sample(cands, n_splits, a.steps + W_)
        for _ in range(W_):
            run_reference_step(w, idx, cores)
        tot, fits = 0.0, 0
        for _ in range(a.steps):
            dt, nf, _ = run_reference_step(w, idx, cores)
            tot += dt
            fits += nf
        v = fits / tot
        emit(({
            "impl": "reference", "metric": "candidate-fits/sec", "value": v, "unit": "fits/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": W_, "ms_per_step": 1e3 * tot / max(a.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": "fits/s", "cores": cores, "kind": "reference", "sample": desc,
                             "what": "scikit-learn %s GridSearchCV(n_jobs=%d, refit=False): the reference's own CPU path "
                                     "(spark_sklearn is not importable here: no pyspark/JVM)" % (__import__("sklearn").__version__, cores)},
            "e2e": {"value": v, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ---------------- our arm -----------------------------------------------------------------------
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from spark_sklearn_b200 import GridSearchCV, RandomizedSearchCV
    from spark_sklearn_b200.estimators import get_engine, adapter_for
    from spark_sklearn_b200.base_search import _dist as D
    from sklearn.base import is_classifier
    from sklearn.model_selection import check_cv
    from spark_sklearn_b200.estimators import fold_ids_from_splits

    est = WL.make_estimator(w)
    X, y = w["X"], w["y"]
    eng = get_engine(local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # resident-data plan (value): dataset uploaded once, the search called K times
    splits = list(check_cv(w["cv"], y, classifier=is_classifier(est)).split(X, y))
    fold_id = fold_ids_from_splits(splits, len(y))
    plan = adapter_for(est).plan(est, cands, X, y, fold_id, len(splits))       # gs_set_data happens here
    parts = D.assign_for_plan(plan, len(cands), world)                                     # same dealing as GridSearchCV.fit
    my = parts[rank]

    def resident_step():
        local = plan.evaluate(my, return_train=True)
        out = D.allgather_candidates(local, my, len(cands), len(splits), world, parts)
        return out, plan.profile()

    for _ in range(W_):
        resident_step()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev_ms = solve_ms = 0.0
    launches = 0
    iters = 0
    sbytes = 0.0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        out, prof = resident_step()
        ev_ms += prof["ms_total"]
        solve_ms += prof["ms_solve"]
        launches += int(prof["launches"])
        iters += int(prof.get("smo_iterations", 0))
        sbytes += float(prof.get("solve_bytes", 0.0))
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    test_scores = out["test"]

    # end to end through the public API with host buffers
    def e2e_step():
        if w["search"] == "grid":
            s = GridSearchCV(None, est, w["param_grid"], cv=w["cv"], refit=False)
        else:
            s = RandomizedSearchCV(None, est, w["param_distributions"], n_iter=w["n_iter"], cv=w["cv"], refit=False,
                                   random_state=w["random_state"])
        s.fit(X, y)
        return s
    for _ in range(min(W_, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        s = e2e_step()
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_prof = s.device_profile_

    # max over ranks
    times = torch.tensor([ev_ms, wall, e2e_wall, solve_ms], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(launches), float(iters), sbytes], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ev_ms, wall, e2e_wall, solve_ms = [float(x) for x in times.cpu()]
    launches, iters, sbytes = [float(x) for x in sums.cpu()]
    fits = len(cands) * len(splits)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    # roofline of the dominant kernel: batched SMO, HBM-bound by design (row gathers)
    per_launch_bytes = sbytes / max(a.steps, 1) / max(world, 1)
    per_launch_s = solve_ms / max(a.steps, 1) * 1e-3
    achieved = per_launch_bytes / per_launch_s / 1e9 if per_launch_s > 0 else 0.0
    # DRAM traffic of the SMO launches of ONE config-2 step, from ncu (profiles/r01_smo_dram_full_c2.csv):
    # smo_kernel 257.32 + 0.87 GB, smo_colown_kernel 30.39 + 0.01 GB.  Only valid for the N=1 config-2 workload.
    traffic = 288.6e9 if (world <= 1 and a.workload == "c2") else None
    roofline = {"kernel": "smo_kernel (one CTA per sub-problem) + smo_colown_kernel (4-CTA cluster per critical-path sub-problem), "
                          "launched concurrently: batched exact-trajectory C-SVC SMO", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": traffic,
                "algorithmic_bytes_per_launch": per_launch_bytes,
                "note": "algorithmic bytes = sum over sub-problems of n_iter * 2 rows * l * 4 B (SURVEY.md 8d) per step (both SMO "
                        "launches); time = the solve phase of the step from CUDA events on the engine stream.  The solver is bound "
                        "by the dependent chain of an SMO iteration and by issue slots, not by HBM bandwidth "
                        "(profiles/r01_smo_final_ncu_summary.txt: 58 % / 33 % issue-active, measured DRAM traffic 0.86x algorithmic)"}
    result = {
        "metric": "candidate-fits/sec", "value": a.steps * fits / (ev_ms * 1e-3), "unit": "fits/s", "n_gpus": max(world, 1),
        "steps": a.steps, "warmup": W_, "ms_per_step": ev_ms / max(a.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "wall_ms_per_step": 1e3 * wall / max(a.steps, 1),
        "e2e": {"value": a.steps * fits / e2e_wall, "unit": "fits/s",
                "h2d_bytes_per_step": int(e2e_prof.get("h2d_bytes", 0)), "d2h_bytes_per_step": int(e2e_prof.get("d2h_bytes", 0)),
                "api": "spark_sklearn_b200.GridSearchCV(sc=None, ..., refit=False).fit(X, y) with host numpy arrays"},
        "gpu_launches": int(launches), "smo_iterations_per_step": iters / max(a.steps, 1),
        "roofline": roofline, "clocks": clocks,
        "best_mean_test_score": float(np.max(np.mean(test_scores, 1))),
    }
    if not a.no_cpu_baseline and world == 1:
        idx, desc = cpu_sample(cands, n_splits, 2)
        dt, nf, cpu_mean = run_reference_step(w, idx, cores)
        gpu_mean = np.mean(test_scores[idx], 1)
        result["cpu_baseline"] = {"value": nf / dt, "unit": "fits/s", "cores": cores, "kind": "reference", "sample": desc,
                                  "seconds": dt,
                                  "max_abs_diff_mean_test_score_vs_gpu": float(np.max(np.abs(cpu_mean - gpu_mean)))}
    emit(result)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
