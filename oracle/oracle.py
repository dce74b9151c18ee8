"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference hot path: the per-(candidate, fold) fit+score task the
reference maps over Spark (reference python/spark_sklearn/base_search.py:56-90), whose
arithmetic is scikit-learn's (``_fit_and_score`` -> SVC / Ridge / LogisticRegression).
"SK/" below = site-packages/sklearn (1.9.0 in this image; third-party dependency of the
reference, pinned there to >=0.18.1,<0.20 -- python/setup.py:22).

* SVC: ``svc_oracle.c`` (C restatement of SK/svm/src/libsvm/svm.cpp) driven from here for
  class grouping / one-vs-one / voting (svm.cpp:2246-2327, 2441-2523, 2821-2904).
* Ridge: numpy restatement of SK/linear_model/_ridge.py:215-227 (+ _base.py:113 centring).
* Lasso / ElasticNet: numpy restatement of SK/linear_model/_cd_fast.pyx:243-506 (cyclic coordinate descent on the
  residual, duality-gap stop, gap-safe screening) with the scaling of _coordinate_descent.py:781-782.
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


# ------------------------------------------------------- Lasso / ElasticNet -------------
def enet_cd(X, y, alpha, beta, tol=1e-4, max_iter=1000, do_screening=True):
    """SK/linear_model/_cd_fast.pyx:243-506 enet_coordinate_descent (dense X, cyclic, positive=False) in the dtype of X.
    alpha / beta = L1 / L2 penalties already scaled by n_samples.  -> (w, gap, tol * y.y, n_iter)"""
    dt = X.dtype.type
    n, d = X.shape
    alpha, beta = dt(alpha), dt(beta)
    norm2 = np.einsum("ij,ij->j", X, X, dtype=X.dtype)                # :371-373
    w = np.zeros(d, X.dtype)
    R = y.copy()                                                         # :403-406 (w = 0)
    d_w_tol = dt(tol)
    tol = dt(tol) * dt(y @ y)                                            # :409

    def gap_enet():                                                      # :162-240
        R_norm2 = dt(R @ R)
        w_l2 = dt(w @ w) if beta > 0 else dt(0)
        Ry = dt(R @ y)
        XtR = (X.T @ R).astype(X.dtype)
        if alpha == 0:
            dn = dt(XtR @ XtR)
            if beta == 0:
                return dn, dn, XtR
            return dt(R_norm2 + dt(0.5) * beta * w_l2 - Ry + dn / (2 * beta)), dn, XtR
        XtA = (XtR - beta * w).astype(X.dtype)
        dn = dt(np.abs(XtA).max())
        primal = dt(0.5) * (R_norm2 + beta * w_l2) + alpha * dt(np.abs(w).sum())      # :138-159
        scale = alpha / dn if dn > alpha else dt(1)
        dual = dt(-0.5) * scale * scale * (R_norm2 + beta * w_l2) + scale * Ry
        return dt(primal - dual), dn, XtA

    screening = do_screening and alpha != 0                              # :391-393
    active = np.arange(d)
    excluded = np.zeros(d, bool)

    def screen(gap, dn, XtA, first):                                     # :399-422, :473-492
        nonlocal active, R
        keep = []
        for j in range(d):
            if first and norm2[j] == 0:
                w[j] = 0
                excluded[j] = True
                continue
            if not first and excluded[j]:
                continue
            with np.errstate(divide="ignore", invalid="ignore"):
                d_j = (1 - abs(XtA[j] / max(alpha, dn))) / np.sqrt(norm2[j] + beta)
            if d_j <= np.sqrt(2 * gap) / alpha:
                keep.append(j)
                excluded[j] = False
            else:
                if w[j] != 0:
                    R += w[j] * X[:, j]
                    w[j] = 0
                excluded[j] = True
        active = np.array(keep, int)

    gap, dn, XtA = gap_enet()                                            # :411-417
    if gap <= tol:
        return w, gap, tol, 0
    if screening:
        screen(gap, dn, XtA, True)
    n_iter = 0
    for n_iter in range(max_iter):                                       # :424
        w_max = d_w_max = dt(0)
        for j in active:
            if norm2[j] == 0:
                continue
            w_j = w[j]
            tmp = dt(X[:, j] @ R) + w_j * norm2[j]                       # :441
            w[j] = np.sign(tmp) * max(abs(tmp) - alpha, 0) / (norm2[j] + beta)
            if w[j] != w_j:
                R += (w_j - w[j]) * X[:, j]                              # :449-451
            d_w_max = max(d_w_max, abs(w[j] - w_j))
            w_max = max(w_max, abs(w[j]))
        if w_max == 0 or d_w_max / w_max <= d_w_tol or n_iter == max_iter - 1:       # :458-462
            gap, dn, XtA = gap_enet()
            if gap <= tol:
                break
            if screening:
                screen(gap, dn, XtA, False)
    return w, gap, tol, n_iter + 1


def enet_fit_score(X, y, train, test, alpha, l1_ratio=1.0, fit_intercept=True, tol=1e-4, max_iter=1000):
    """SK/linear_model/_coordinate_descent.py:1170-1280 (ElasticNet.fit: centre by the training means, penalties scaled by
    n_samples :781-782), SK/base.py:716 (r2).  -> (test r2, train r2, n_iter)"""
    dt = X.dtype
    Xt, yt = X[train], y[train].astype(dt)
    if fit_intercept:
        xm = Xt.mean(0, dtype=np.float64).astype(dt)
        ym = yt.mean(dtype=np.float64).astype(dt)
        Xc, yc = Xt - xm, yt - ym
    else:
        Xc, yc = Xt, yt
    n = len(train)
    w, _gap, _tol, n_iter = enet_cd(np.asfortranarray(Xc), yc, alpha * l1_ratio * n, alpha * (1.0 - l1_ratio) * n, tol, max_iter)
    b = (ym - xm @ w) if fit_intercept else dt.type(0)

    def r2(rows):
        yp = X[rows] @ w + b
        yt_ = y[rows].astype(np.float64)
        return 1.0 - ((yt_ - yp.astype(np.float64)) ** 2).sum() / ((yt_ - yt_.mean()) ** 2).sum()
    return r2(test), r2(train), n_iter


def cv_scores_enet(X, y, fold_id, n_splits, cands, fit_intercept=True):
    allrows = np.arange(len(y))
    test = np.zeros((len(cands), n_splits))
    train = np.zeros_like(test)
    iters = np.zeros(test.shape, int)
    for ci, p in enumerate(cands):
        for k in range(n_splits):
            test[ci, k], train[ci, k], iters[ci, k] = enet_fit_score(
                X, y, allrows[fold_id != k], allrows[fold_id == k], p.get("alpha", 1.0), p.get("l1_ratio", 1.0), fit_intercept,
                p.get("tol", 1e-4), p.get("max_iter", 1000))
    return test, train, iters


# ------------------------------------------------------- LogisticRegression ------------
def logreg_fit_score(X, y, train, test, C, tol=1e-4, max_iter=100, fit_intercept=True):
    """L2 logistic regression, lbfgs: SK/linear_model/_logistic.py:403 (dtype kept),
    :580-604 (scipy L-BFGS-B, maxcor default 10, maxls=50, gtol=tol, ftol=64*eps),
    objective SK/linear_model/_linear_loss.py:47-64 with l2_reg_strength = 1/(C*n).
    Two classes:  f(w,b) = (1/n) sum_i [log(1+exp(z_i)) - y_i z_i] + 0.5*l2*|w|^2,   z = Xw + b.
    Three or more (multinomial, one weight row per class; SK/_loss/_loss.pyx closs_grad_half_multinomial):
                  f(W,b) = (1/n) sum_i [logsumexp(z_i) - z_i[y_i]] + 0.5*l2*|W|_F^2,   z_i = W x_i + b."""
    from scipy import optimize
    from scipy.special import logsumexp
    Xt = X[train]
    dt = Xt.dtype
    classes = np.unique(y[train])
    n, d = Xt.shape
    l2 = 1.0 / (C * n)
    opts = {"maxiter": max_iter, "maxls": 50, "gtol": tol, "ftol": 64 * np.finfo(float).eps}
    if len(classes) > 2:
        # float32 X: scikit-learn casts the weights to X.dtype, keeps raw predictions, pointwise losses and gradients in that
        # dtype and contracts them with float32 BLAS (_linear_loss.py:216-222, 329-376); the optimiser state is float64 and
        # the variables are ordered class-fastest (coef.reshape((n_classes, -1), order="F")).
        K = len(classes)
        Y = y[train][:, None] == classes[None, :]
        nv = d + (1 if fit_intercept else 0)

        def fgm(w):
            Wm = w.reshape((K, nv), order="F")
            Z = Xt @ Wm[:, :d].astype(dt).T
            if fit_intercept:
                Z = Z + Wm[:, d].astype(dt)
            Z64 = Z.astype(np.float64)
            mx = Z64.max(1)
            E = np.exp(Z64 - mx[:, None])
            se = E.sum(1)
            loss_i = (np.log(se) + mx - Z64[Y]).astype(dt)
            P = (E / se[:, None] - Y).astype(dt)
            loss = float(loss_i.sum() / n) + 0.5 * l2 * float((Wm[:, :d] * Wm[:, :d]).sum())
            P /= dt.type(n)
            g = np.empty((K, nv), order="F")
            g[:, :d] = (P.T @ Xt) + l2 * Wm[:, :d]
            if fit_intercept:
                g[:, d] = P.sum(0)
            return loss, g.ravel(order="F")
        res = optimize.minimize(fgm, np.zeros(K * nv), method="L-BFGS-B", jac=True, options=opts)
        Wm = res.x.reshape((K, nv), order="F")

        def accm(rows):
            Z = X[rows].astype(np.float64) @ Wm[:, :d].T + (Wm[:, d] if fit_intercept else 0.0)
            return np.mean(classes[Z.argmax(1)] == y[rows])
        return accm(test), accm(train), res.nit
    yt = (y[train] == classes[1]).astype(dt)

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
    res = optimize.minimize(fg, w0, method="L-BFGS-B", jac=True, options=opts)
    w = res.x

    def acc(rows):
        z = X[rows].astype(np.float64) @ w[:d] + (w[d] if fit_intercept else 0.0)
        return np.mean(classes[(z > 0).astype(int)] == y[rows])
    return acc(test), acc(train), res.nit


def cv_scores_logreg(X, y, fold_id, n_splits, cands, return_n_iter=False):
    allrows = np.arange(len(y))
    test = np.zeros((len(cands), n_splits))
    train = np.zeros_like(test)
    iters = np.zeros(test.shape, int)
    for ci, p in enumerate(cands):
        for k in range(n_splits):
            test[ci, k], train[ci, k], iters[ci, k] = logreg_fit_score(X, y, allrows[fold_id != k],
                                                                       allrows[fold_id == k], p.get("C", 1.0))
    return (test, train, iters) if return_n_iter else (test, train)
