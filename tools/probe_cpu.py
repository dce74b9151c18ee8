import os, time, numpy as np
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    try: print(p, open(p).read().strip())
    except Exception as e: pass
import sys; sys.path.insert(0, '.')
from spark_sklearn_b200 import workloads as W
from sklearn.svm import SVC
w = W.make_workload("c2"); X, y = w["X"], w["y"]
tr = np.arange(8000)
t=time.time(); s=SVC(C=1.0, gamma=1/512).fit(X[tr], y[tr]); print("one fit C=1 g=1/512: %.1fs n_iter %d" % (time.time()-t, s.n_iter_[0]))
from joblib import Parallel, delayed
def f(i):
    t=time.time(); SVC(C=1.0, gamma=1/512).fit(X[tr], y[tr]); return time.time()-t
for nj in (8, 32, 64):
    t=time.time(); r=Parallel(n_jobs=nj)(delayed(f)(i) for i in range(nj)); print("n_jobs", nj, "wall %.1f" % (time.time()-t), "per-fit mean %.1f max %.1f" % (np.mean(r), np.max(r)))
