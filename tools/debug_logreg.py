import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spark_sklearn_b200 import workloads as W
from spark_sklearn_b200.estimators import get_engine
from sklearn.linear_model import LogisticRegression
w = W.make_workload("c3_small"); X, y = w["X"], w["y"]
eng = get_engine(0)
eng.set_data(X, np.full(len(y), -1, np.int8), 1, y_class=y.astype(np.int32))
for C in (1e-3, 1e-1, 10.0):
    wg, bg, it = eng.logreg_refit(C)
    s = LogisticRegression(C=C).fit(X, y)
    ws, bs = s.coef_[0], s.intercept_[0]
    z_g = X.astype(np.float64) @ wg + bg; z_s = X.astype(np.float64) @ ws + bs
    def obj(wv, b):
        z = X.astype(np.float64) @ wv + b
        return np.mean(np.logaddexp(0, z) - y * z) + 0.5 / (C * len(y)) * wv @ wv
    print("C=%g  n_iter gpu %d sk %d | coef rel diff %.3e | intercept %.6f vs %.6f | obj gpu %.12f sk %.12f | pred flips %d"
          % (C, it, s.n_iter_[0], np.abs(wg - ws).max() / np.abs(ws).max(), bg, bs, obj(wg, bg), obj(ws, bs), ((z_g > 0) != (z_s > 0)).sum()))
    print("   first coefs gpu", wg[:4], " sk", ws[:4])
