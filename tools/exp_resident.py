"""Development: does the step time depend on re-uploading the data, or on nvidia-smi polling during the run?"""
import sys, os, time, subprocess, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spark_sklearn_b200 import workloads as W
from spark_sklearn_b200.estimators import get_engine, fold_ids_from_splits
from sklearn.model_selection import StratifiedKFold
w = W.make_workload("c2"); X, y = w["X"], w["y"]
eng = get_engine(0)
fold_id = fold_ids_from_splits(list(StratifiedKFold(5).split(X, y)), len(y))
cands = W.candidates(w)
Cs = [c["C"] for c in cands]; G = np.array([c["gamma"] for c in cands])[:, None]
def run(tag, n, setdata):
    ts = []
    for _ in range(n):
        if setdata: eng.set_data(X, fold_id, 5, y_class=y.astype(np.int32))
        eng.svc(["rbf"] * len(cands), Cs, G); p = eng.profile(); ts.append({k[3:]: round(v, 1) for k, v in p.items() if k.startswith("ms_")})
    print(tag, ts, flush=True)
eng.set_data(X, fold_id, 5, y_class=y.astype(np.int32))
run("warm", 2, False)
run("resident (no set_data)", 4, False)
run("set_data each step", 4, True)
