import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spark_sklearn_b200 import workloads as W
from spark_sklearn_b200.estimators import get_engine, fold_ids_from_splits
from sklearn.model_selection import StratifiedKFold
w = W.make_workload("c2"); X, y = w["X"], w["y"]
eng = get_engine(0)
fold_id = fold_ids_from_splits(list(StratifiedKFold(5).split(X, y)), len(y))
eng.set_data(X, fold_id, 5, y_class=y.astype(np.int32))
C = [10.0, 31.6]; G = [1/1024, 1/1024]
r = eng.svc(["rbf"]*2, C, np.array(G)[:, None]); p = eng.profile()
print("iters", r["n_iter"].sum(), "solve ms", p["ms_solve"], "us/iter", (r["fit_ms"]*1e3/r["n_iter"]).round(2))
