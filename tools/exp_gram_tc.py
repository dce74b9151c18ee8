import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spark_sklearn_b200 import workloads as W
from spark_sklearn_b200.estimators import get_engine, fold_ids_from_splits
from spark_sklearn_b200.engine import GS_GRAM_TENSOR
from sklearn.model_selection import StratifiedKFold
w = W.make_workload("c2"); X, y = w["X"], w["y"]
eng = get_engine(0)
fold_id = fold_ids_from_splits(list(StratifiedKFold(5).split(X, y)), len(y))
eng.set_data(X, fold_id, 5, y_class=y.astype(np.int32))
for fl in (0, GS_GRAM_TENSOR, GS_GRAM_TENSOR):
    r = eng.svc(["rbf"], [0.1], [[1/256]], flags=fl); p = eng.profile()
    print("flags", fl, "gram ms %.3f" % p["ms_gram"], "test", r["test"].round(4), "iters", r["n_iter"])
