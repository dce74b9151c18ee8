import sys, os, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spark_sklearn_b200 import workloads as W
from spark_sklearn_b200.estimators import get_engine, fold_ids_from_splits
from sklearn.model_selection import StratifiedKFold
w = W.make_workload("c2"); X, y = w["X"], w["y"]
eng = get_engine(0)
fold_id = fold_ids_from_splits(list(StratifiedKFold(5).split(X, y)), len(y))
eng.set_data(X, fold_id, 5, y_class=y.astype(np.int32))
Cs = np.logspace(-1, 2.5, 8); gs = np.geomspace(1/4096, 1/256, 8)
def run(name, Cl, gl):
    C = [c for c in Cl for g in gl]; G = [g for c in Cl for g in gl]
    for rep in range(2):
        r = eng.svc(["rbf"]*len(C), C, np.array(G)[:, None]); p = eng.profile()
    us = r["fit_ms"]*1e3/r["n_iter"]
    print("%-28s problems %3d solve %.1f ms | us/iter median %.2f min %.2f max %.2f | iters max %d" % (name, len(C)*5, p["ms_solve"], np.median(us), us.min(), us.max(), r["n_iter"].max()), flush=True)
run("1 gamma x 8 C (40 probs)", Cs, gs[:1])
run("1 gamma(last) x 8 C", Cs, gs[-1:])
run("8 gamma x 1 C=316 (40)", Cs[-1:], gs)
run("2 gamma x 8 C (80)", Cs, gs[:2])
run("4 gamma x 8 C (160)", Cs, gs[:4])
run("8 gamma x 8 C (320)", Cs, gs)
