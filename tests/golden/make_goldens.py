#!/usr/bin/env python
"""Generate golden cv_results_ score arrays with the CPU oracle (scikit-learn 1.9.0).

TEST INFRASTRUCTURE.  The reference (databricks/spark-sklearn) cannot be imported in
this image (no pyspark/JVM; pinned to scikit-learn <0.20), and its own
tests pin no numeric result on this path.  Every floating-point operation of the
reference's hot path happens inside scikit-learn's ``_fit_and_score`` (reference
``base_search.py:83-87``), so the oracle is scikit-learn itself:

* mode "search": ``sklearn.model_selection.GridSearchCV/RandomizedSearchCV`` run as a whole
  (what the reference README now recommends) -- used for the small configs;
* mode "tasks":  the reference's own task list ``(candidate, fold)`` (``base_search.py:56-61``)
  mapped with joblib over a restatement of ``fun`` (``base_search.py:74-88``): clone,
  ``X[train]``, ``fit``, ``score(test)``, ``score(train)`` -- plus solver diagnostics
  (``n_iter_``, nSV, margin-crowding counts) that ``cv_results_`` does not expose.
  Used for the full-size configs (minutes to an hour of CPU).  tests/test_oracle.py checks
  that both modes give identical per-split scores.

Usage:  python tests/golden/make_goldens.py c1 c2_small ... [--jobs N] [--mode tasks|search]
Writes tests/golden/<workload-name>.npz
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from spark_sklearn_b200 import workloads as W  # noqa: E402


def _one_task(est, X, y, params, train, test):
    """Restatement of the reference's per-task closure (base_search.py:74-88)."""
    from sklearn.base import clone
    e = clone(est).set_params(**params)
    t0 = time.time()
    e.fit(X[train], y[train])
    t1 = time.time()
    te = e.score(X[test], y[test])
    t2 = time.time()
    tr = e.score(X[train], y[train])
    diag = np.zeros(4)
    n_iter = getattr(e, "n_iter_", 0)
    diag[0] = float(np.sum(n_iter)) if n_iter is not None else 0.0
    if hasattr(e, "n_support_"):
        diag[1] = float(np.sum(e.n_support_))
        if len(e.classes_) == 2:
            dec = e.decision_function(X[test])
            diag[2] = float(np.sum(np.abs(dec) < 1e-3))   # test points inside libsvm's own tolerance
            diag[3] = float(np.sum(np.abs(dec) < 1e-6))
    return te, tr, len(test), t1 - t0, t2 - t1, diag


def run(key, jobs, mode):
    from joblib import Parallel, delayed
    from sklearn.base import is_classifier
    from sklearn.model_selection import check_cv
    w = W.make_workload(key)
    X, y = w["X"], w["y"]
    est = W.make_estimator(w)
    cands = W.candidates(w)
    cv = check_cv(w["cv"], y, classifier=is_classifier(est))
    splits = list(cv.split(X, y))
    n_splits = len(splits)
    fold_id = np.full(len(y), -1, np.int8)
    for k, (_, te) in enumerate(splits):
        fold_id[te] = k
    t0 = time.time()
    out = {}
    if mode == "search":
        from sklearn.model_selection import GridSearchCV, RandomizedSearchCV
        if w["search"] == "grid":
            s = GridSearchCV(est, w["param_grid"], cv=w["cv"], return_train_score=True, n_jobs=jobs)
        else:
            s = RandomizedSearchCV(est, w["param_distributions"], n_iter=w["n_iter"], cv=w["cv"],
                                   random_state=w["random_state"], return_train_score=True, n_jobs=jobs)
        s.fit(X, y)
        r = s.cv_results_
        test = np.stack([r["split%d_test_score" % k] for k in range(n_splits)], 1)
        train = np.stack([r["split%d_train_score" % k] for k in range(n_splits)], 1)
        diag = np.zeros((len(cands), n_splits, 4))
        assert [dict(p) for p in r["params"]] == [dict(p) for p in cands]
        out["best_index"] = s.best_index_
    else:
        res = Parallel(n_jobs=jobs, verbose=1)(
            delayed(_one_task)(est, X, y, p, tr, te) for p in cands for (tr, te) in splits)
        test = np.array([r[0] for r in res]).reshape(len(cands), n_splits)
        train = np.array([r[1] for r in res]).reshape(len(cands), n_splits)
        diag = np.array([r[5] for r in res]).reshape(len(cands), n_splits, 4)
        out["fit_time"] = np.array([r[3] for r in res]).reshape(len(cands), n_splits)
        out["score_time"] = np.array([r[4] for r in res]).reshape(len(cands), n_splits)
    wall = time.time() - t0
    keys = sorted(cands[0].keys())
    out.update(dict(
        test_scores=test, train_scores=train, diag=diag, fold_id=fold_id,
        mean_test_score=test.mean(1), mean_train_score=train.mean(1),
        param_names=np.array(keys), param_values=np.array([[repr(p[k]) for k in keys] for p in cands]),
        wall_s=wall, jobs=jobs, mode=mode, sklearn_version=__import__("sklearn").__version__,
        x_checksum=float(np.asarray(X, np.float64).sum()),
    ))
    path = os.path.join(ROOT, "tests", "golden", w["name"] + ".npz")
    np.savez_compressed(path, **out)
    print("%s: %d cand x %d folds in %.1fs (%d jobs, mode=%s) -> %s  best mean_test=%.4f"
          % (key, len(cands), n_splits, wall, jobs, mode, path, test.mean(1).max()), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("keys", nargs="+")
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--mode", default="tasks")
    a = ap.parse_args()
    for k in a.keys:
        run(k, a.jobs, a.mode)


def make_refit_golden():
    """Full-data fit of one config-2 candidate (the refit path): tests/golden/c2_refit_C10_g1024.npz."""
    import numpy as np
    from sklearn.svm import SVC
    from spark_sklearn_b200 import workloads as W
    w = W.make_workload("c2")
    X, y = w["X"], w["y"]
    s = SVC(kernel="rbf", C=10.0, gamma=1 / 1024).fit(X, y)
    from oracle import oracle as O                         # the C restatement (Gram formed like the GPU's): bit-level target
    rows = np.concatenate([np.flatnonzero(y == 0), np.flatnonzero(y == 1)]).astype(np.int32)
    oc, orho, oit, _ = O.svc_solve(X.astype(np.float64), rows, int((y == 0).sum()), "rbf", 1 / 1024, 10.0)
    full = np.zeros(len(y)); full[rows] = oc
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c2_refit_C10_g1024.npz"), n_iter=s.n_iter_,
                        n_support=s.n_support_, intercept=s.intercept_, support=s.support_.astype(np.int32), dual_coef=s.dual_coef_,
                        oracle_coef=full, oracle_rho=np.array([orho]))
