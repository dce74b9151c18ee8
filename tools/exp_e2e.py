"""Where does a GridSearchCV.fit() on config 2 spend its time?  (e2e vs resident gap)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200GS_DEVICES", "1")
import numpy as np
from spark_sklearn_b200 import GridSearchCV, workloads as W
from spark_sklearn_b200.estimators import get_engine
key = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = W.make_workload(key)
est = W.make_estimator(w)
def make():
    if "param_distributions" in w:
        from spark_sklearn_b200 import RandomizedSearchCV
        return RandomizedSearchCV(None, est, w["param_distributions"], n_iter=w["n_iter"], cv=w["cv"], refit=False, random_state=w.get("random_state", 0))
    return GridSearchCV(None, est, w["param_grid"], cv=w["cv"], refit=False)
for rep in range(4):
    t0 = time.perf_counter()
    s = make().fit(w["X"], w["y"])
    dt = time.perf_counter() - t0
    p = s.device_profile_
    print("fit %d: wall %.1f ms | device total %.1f solve %.1f gram %.1f kmat %.1f score %.1f h2d %.1f" % (
        rep, dt * 1e3, p["ms_total"], p["ms_solve"], p["ms_gram"], p["ms_kernel_matrix"], p["ms_score"], p["ms_h2d"]), flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
make().fit(w["X"], w["y"])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
