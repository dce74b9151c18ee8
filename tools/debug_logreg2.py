import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spark_sklearn_b200 import workloads as W
from spark_sklearn_b200.estimators import get_engine
from sklearn.linear_model import LogisticRegression
from sklearn.model_selection import StratifiedKFold
w = W.make_workload("c3_small"); X, y = w["X"], w["y"]
eng = get_engine(0)
splits = list(StratifiedKFold(5).split(X, y))
fold_id = np.zeros(len(y), np.int8)
for k, (_, te) in enumerate(splits): fold_id[te] = k
Cs = [1e-3, 0.1, 10.0, 50.0]
eng.set_data(X, fold_id, 5, y_class=y.astype(np.int32))
r = eng.logreg(Cs)
for ci, C in enumerate(Cs):
    for k in (0, 3):
        tr, te = splits[k]
        s = LogisticRegression(C=C).fit(X[tr], y[tr])
        eng.set_data(X[tr], np.full(len(tr), -1, np.int8), 1, y_class=y[tr].astype(np.int32))
        wg, bg, it = eng.logreg_refit(C)
        zt = X[te].astype(np.float64) @ wg + bg
        zr = X[tr].astype(np.float64) @ wg + bg
        print("C=%g fold %d: CV-column test %.5f train %.5f it %d | sklearn test %.5f train %.5f it %d | refit-on-subset test %.5f train %.5f it %d  coef rel diff vs sk %.2e"
              % (C, k, r["test"][ci, k], r["train"][ci, k], r["n_iter"][ci, k], s.score(X[te], y[te]), s.score(X[tr], y[tr]), s.n_iter_[0],
                 np.mean((zt > 0) == y[te]), np.mean((zr > 0) == y[tr]), it, np.abs(wg - s.coef_[0]).max() / np.abs(s.coef_[0]).max()))
