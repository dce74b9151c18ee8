"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference hot path: the per-(candidate, fold) fit+score task the
reference maps over Spark (reference python/spark_sklearn/base_search.py:56-90), whose
arithmetic is scikit-learn's (``_fit_and_score`` -> SVC / Ridge / LogisticRegression).
"SK/" below = site-packages/sklearn (1.9.0 in this image; third-party dependency of the
reference, pinned there to >=0.18.1,<0.20 -- python/setup.py:22).

* SVC: ``svc_oracle.c`` (C restatement of SK/svm/src/libsvm/svm.cpp) driven from here for
  class grouping / one-vs-one / voting (svm.cpp:2246-2327, 2441-2523, 2821-2904).
* Ridge: numpy restatement of SK/linear_model/_ridge.py:215-227 (+ _base.py:113 centring).
* LogisticRegression: scipy L-BFGS-B on the restated objective of
  SK/linear_model/_linear_loss.py:47-64 with the options of _logistic.py:585-597.

Pinned by tests/test_oracle.py against sklearn itself and tests/golden/*.npz.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
KERNEL_ID = {"linear": 0, "rbf": 1}


def build(force=False):
    """gcc -O2 -ffp-contract=off -shared svc_oracle.c -> libsvc_oracle.so (next to this file)."""
    so = os.path.join(_HERE, "libsvc_oracle.so")
    src = os.path.join(_HERE, "svc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        dp, ip, fp = (ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int),
                      ctypes.POINTER(ctypes.c_float))
        L.oracle_svc_solve.argtypes = [dp, ctypes.c_int, ctypes.c_int, ip, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_int, ctypes.c_int, fp, ctypes.c_long,
                                       dp, dp, ip, dp]
        L.oracle_svc_solve.restype = ctypes.c_int
        L.oracle_svc_decision.argtypes = [dp, ctypes.c_int, ip, ctypes.c_int, dp, ctypes.c_double,
                                          ctypes.c_int, ctypes.c_double, ip, ctypes.c_int, dp]
        L.oracle_svc_decision.restype = None
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def svc_solve(X64, rows, n_pos, kernel, gamma, C, tol=1e-3, shrinking=True, max_iter=-1, Kpre=None):
    """One binary sub-problem.  Returns (coef[l] = alpha*y, rho, n_iter, obj)."""
    L = _lib()
    X64 = np.ascontiguousarray(X64, np.float64)
    rows = np.ascontiguousarray(rows, np.int32)
    l = len(rows)
    coef = np.zeros(l)
    rho, obj, it = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    kp, ldk = None, 0
    if Kpre is not None:
        Kpre = np.ascontiguousarray(Kpre, np.float32)
        kp, ldk = _p(Kpre, ctypes.c_float), Kpre.shape[1]
    L.oracle_svc_solve(_p(X64, ctypes.c_double), X64.shape[0], X64.shape[1], _p(rows, ctypes.c_int), l,
                       int(n_pos), KERNEL_ID[kernel], float(gamma), float(C), float(tol), int(shrinking),
                       int(max_iter), kp, ldk, _p(coef, ctypes.c_double), ctypes.byref(rho),
                       ctypes.byref(it), ctypes.byref(obj))
    return coef, rho.value, it.value, obj.value


def svc_decision(X64, rows, coef, rho, kernel, gamma, trows):
    L = _lib()
    rows = np.ascontiguousarray(rows, np.int32)
    trows = np.ascontiguousarray(trows, np.int32)
    coef = np.ascontiguousarray(coef, np.float64)
    out = np.zeros(len(trows))
    L.oracle_svc_decision(_p(X64, ctypes.c_double), X64.shape[1], _p(rows, ctypes.c_int), len(rows),
                          _p(coef, ctypes.c_double), float(rho), KERNEL_ID[kernel], float(gamma),
                          _p(trows, ctypes.c_int), len(trows), _p(out, ctypes.c_double))
    return out


def resolve_gamma(gamma, Xtrain64):
    """SK/svm/_base.py:278-286."""
    if isinstance(gamma, str):
        if gamma == "scale":
            v = Xtrain64.var()
            return 1.0 / (Xtrain64.shape[1] * v) if v != 0 else 1.0
        if gamma == "auto":
            return 1.0 / Xtrain64.shape[1]
        raise ValueError(gamma)
    return float(gamma)


class SVCModel:
    """Fitted one-vs-one C-SVC (svm.cpp:2441-2523 training loop, :2821-2904 prediction)."""

    def __init__(self, X64, y, train, kernel="rbf", gamma="scale", C=1.0, tol=1e-3, shrinking=True,
                 max_iter=-1):
        self.X64 = X64
        train = np.asarray(train)
        self.classes = np.unique(y[train])              # sorted labels (sklearn's svm_group_classes)
        self.kernel = kernel
        self.gamma = resolve_gamma(gamma, X64[train])
        self.pairs = []
        self.n_iter = []
        by_class = [train[y[train] == c] for c in self.classes]     # original order within class
        for a in range(len(self.classes)):
            for b in range(a + 1, len(self.classes)):
                rows = np.concatenate([by_class[a], by_class[b]]).astype(np.int32)
                coef, rho, it, obj = svc_solve(X64, rows, len(by_class[a]), kernel, self.gamma, C, tol,
                                               shrinking, max_iter)
                self.pairs.append((a, b, rows, coef, rho))
                self.n_iter.append(it)

    def decision_pairs(self, rows):
        return np.stack([svc_decision(self.X64, r, coef, rho, self.kernel, self.gamma, rows)
                         for (_, _, r, coef, rho) in self.pairs], 1)

    def predict(self, rows):
        dec = self.decision_pairs(rows)
        votes = np.zeros((len(rows), len(self.classes)), np.int64)
        for p, (a, b, _, _, _) in enumerate(self.pairs):
            pos = dec[:, p] > 0
            votes[pos, a] += 1
            votes[~pos, b] += 1
        return self.classes[np.argmax(votes, 1)]          # first maximum wins (svm.cpp:2889-2892)


def folds_from_cv(cv, X, y, classifier):
    """fold_id[n]: index of the split whose TEST set holds the row (reference base_search.py:34,
    check_cv -> (Stratified)KFold, SK/model_selection/_split.py:437,774-842)."""
    from sklearn.model_selection import check_cv
    cvo = check_cv(cv, y, classifier=classifier)
    fold_id = np.full(len(y), -1, np.int8)
    n_splits = 0
    for k, (_, te) in enumerate(cvo.split(X, y)):
        fold_id[te] = k
        n_splits = k + 1
    return fold_id, n_splits


def cv_scores_svc(X, y, fold_id, n_splits, cands, est_params=None):
    """[n_cand, n_splits] test / train accuracy + n_iter, candidate-major, fold-minor
    (reference base_search.py:56-61 task order)."""
    est_params = dict(est_params or {})
    X64 = np.ascontiguousarray(X, np.float64)
    allrows = np.arange(len(y), dtype=np.int32)
    test = np.zeros((len(cands), n_splits))
    train = np.zeros_like(test)
    iters = np.zeros((len(cands), n_splits), np.int64)
    for ci, p in enumerate(cands):
        prm = dict(kernel="rbf", gamma="scale", C=1.0, tol=1e-3, shrinking=True, max_iter=-1)
        prm.update({k: v for k, v in est_params.items() if k in prm})
        prm.update({k: v for k, v in p.items() if k in prm})
        for k in range(n_splits):
            tr, te = allrows[fold_id != k], allrows[fold_id == k]
            m = SVCModel(X64, y, tr, **prm)
            test[ci, k] = np.mean(m.predict(te) == y[te])
            train[ci, k] = np.mean(m.predict(tr) == y[tr])
            iters[ci, k] = sum(m.n_iter)
    return test, train, iters


# ---------------------------------------------------------------- Ridge ---------------
def ridge_fit_score(X, y, train, test, alpha, fit_intercept=True):
    """SK/linear_model/_ridge.py:919-1010 (fit), :215-227 (_solve_cholesky), SK/base.py:716 (r2).
    float32 in -> float32 throughout (ridge keeps X.dtype, _ridge.py:1258)."""
    from scipy import linalg
    dt = X.dtype
    Xt, yt = X[train], y[train].astype(dt)
    if fit_intercept:
        xm = Xt.mean(0, dtype=np.float64).astype(dt)     # _preprocess_data: np.average in float64, cast back
        ym = yt.mean(dtype=np.float64).astype(dt)
        Xc, yc = Xt - xm, yt - ym
    else:
        Xc, yc = Xt, yt
    A = Xc.T @ Xc
    A[np.diag_indices_from(A)] += dt.type(alpha)
    w = linalg.solve(A, Xc.T @ yc, assume_a="pos")
    b = (ym - xm @ w) if fit_intercept else dt.type(0)

    def r2(rows):
        yp = X[rows] @ w + b
        yt_ = y[rows].astype(np.float64)
        num = ((yt_ - yp.astype(np.float64)) ** 2).sum()
        den = ((yt_ - yt_.mean()) ** 2).sum()
        return 1.0 - num / den
    return r2(test), r2(train)


def cv_scores_ridge(X, y, fold_id, n_splits, cands, fit_intercept=True):
    allrows = np.arange(len(y))
    test = np.zeros((len(cands), n_splits))
    train = np.zeros_like(test)
    for ci, p in enumerate(cands):
        for k in range(n_splits):
            test[ci, k], train[ci, k] = ridge_fit_score(X, y, allrows[fold_id != k], allrows[fold_id == k],
                                                        p.get("alpha", 1.0), fit_intercept)
    return test, train


# ------------------------------------------------------- LogisticRegression ------------
def logreg_fit_score(X, y, train, test, C, tol=1e-4, max_iter=100, fit_intercept=True):
    """Binary L2 logistic regression, lbfgs: SK/linear_model/_logistic.py:403 (dtype kept),
    :580-604 (scipy L-BFGS-B, maxcor default 10, maxls=50, gtol=tol, ftol=64*eps),
    objective SK/linear_model/_linear_loss.py:47-64 with l2_reg_strength = 1/(C*n):
        f(w,b) = (1/n) sum_i [log(1+exp(z_i)) - y_i z_i] + 0.5*l2*|w|^2,   z = Xw + b."""
    from scipy import optimize
    Xt = X[train]
    dt = Xt.dtype
    classes = np.unique(y[train])
    yt = (y[train] == classes[1]).astype(dt)
    n, d = Xt.shape
    l2 = 1.0 / (C * n)

    def fg(w):
        wv = w[:d].astype(dt)
        z = Xt @ wv + (dt.type(w[d]) if fit_intercept else dt.type(0))
        z64 = z.astype(np.float64)
        loss = np.mean(np.logaddexp(0.0, z64) - yt * z64) + 0.5 * l2 * float(w[:d] @ w[:d])
        r = ((1.0 / (1.0 + np.exp(-z64))) - yt) / n
        g = np.empty_like(w)
        g[:d] = Xt.T.astype(np.float64) @ r + l2 * w[:d]
        if fit_intercept:
            g[d] = r.sum()
        return loss, g
    w0 = np.zeros(d + 1 if fit_intercept else d)
    res = optimize.minimize(fg, w0, method="L-BFGS-B", jac=True,
                            options={"maxiter": max_iter, "maxls": 50, "gtol": tol,
                                     "ftol": 64 * np.finfo(float).eps})
    w = res.x

    def acc(rows):
        z = X[rows].astype(np.float64) @ w[:d] + (w[d] if fit_intercept else 0.0)
        return np.mean(classes[(z > 0).astype(int)] == y[rows])
    return acc(test), acc(train), res.nit


def cv_scores_logreg(X, y, fold_id, n_splits, cands):
    allrows = np.arange(len(y))
    test = np.zeros((len(cands), n_splits))
    train = np.zeros_like(test)
    for ci, p in enumerate(cands):
        for k in range(n_splits):
            test[ci, k], train[ci, k], _ = logreg_fit_score(X, y, allrows[fold_id != k],
                                                            allrows[fold_id == k], p.get("C", 1.0))
    return test, train
