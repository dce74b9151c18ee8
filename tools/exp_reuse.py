"""Development study: how often does the SMO working set re-use recently fetched kernel rows?
Runs the CPU oracle on one config-2 sub-problem with a trace of (i, j) and reports LRU hit rates."""
import sys, ctypes, time
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as O
from spark_sklearn_b200 import workloads as W
from collections import OrderedDict

key = sys.argv[1] if len(sys.argv) > 1 else "c2"
which = sys.argv[2] if len(sys.argv) > 2 else "longest"
w = W.make_workload(key)
g = np.load("tests/golden/%s.npz" % w["name"], allow_pickle=True)
it = g["diag"][:, :, 0]
c, f = np.unravel_index(np.argmax(it) if which == "longest" else np.argsort(it.ravel())[it.size // 2], it.shape)
cand = W.candidates(w)[c]
print("candidate", c, cand, "fold", f, "n_iter", it[c, f])
fold_id, ns = O.folds_from_cv(w["cv"], w["X"], w["y"], True)
X64 = w["X"].astype(np.float64)
tr = np.flatnonzero(fold_id != f)
y = w["y"][tr]
rows = np.concatenate([tr[y == 0], tr[y == 1]]).astype(np.int32)
lib = O._lib()
cap = 2 * (int(it[c, f]) + 1000)
buf = np.zeros(cap, np.int32)
lib.oracle_svc_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_long]
lib.oracle_svc_trace_len.restype = ctypes.c_long
lib.oracle_svc_set_trace(buf.ctypes.data, cap)
t0 = time.time()
r = O.svc_solve(X64, rows, int((y == 0).sum()), "rbf", float(cand["gamma"]), float(cand["C"]))
n = lib.oracle_svc_trace_len()
lib.oracle_svc_set_trace(None, 0)
print("solved in %.1fs, trace %d pairs" % (time.time() - t0, n // 2))
tr_ = buf[:n].reshape(-1, 2)
np.save("/tmp/trace_%s_%s.npy" % (key, which), tr_)
for capn in (1, 2, 3, 4, 6, 8, 16, 32, 64, 128):
    lru = OrderedDict(); hi = hj = 0
    for i, j in tr_:
        for kk, r_ in enumerate((i, j)):
            if r_ in lru:
                lru.move_to_end(r_)
                if kk == 0: hi += 1
                else: hj += 1
            else:
                lru[r_] = 1
                if len(lru) > capn: lru.popitem(last=False)
    print("LRU %3d rows: hit i %.3f  j %.3f" % (capn, hi / len(tr_), hj / len(tr_)))
# is next i one of {prev i, prev j, prev runner-up}?
same_i = np.mean(tr_[1:, 0] == tr_[:-1, 0]); i_is_prev_j = np.mean(tr_[1:, 0] == tr_[:-1, 1]); j_is_prev_i = np.mean(tr_[1:, 1] == tr_[:-1, 0]); same_j = np.mean(tr_[1:, 1] == tr_[:-1, 1])
print("next i == prev i %.3f, == prev j %.3f; next j == prev i %.3f, == prev j %.3f" % (same_i, i_is_prev_j, j_is_prev_i, same_j))
print("distinct rows %d of %d fetches" % (len(np.unique(tr_)), tr_.size))
