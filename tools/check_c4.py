"""Development: where do config-4 trajectories differ from scikit-learn's (n_iter by candidate / fold)?"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spark_sklearn_b200 import workloads as W
from spark_sklearn_b200.estimators import get_engine, fold_ids_from_splits
from sklearn.model_selection import StratifiedKFold
key = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = W.make_workload(key); X, y = w["X"], w["y"]
eng = get_engine(0)
fold_id = fold_ids_from_splits(list(StratifiedKFold(5).split(X, y)), len(y))
eng.set_data(X, fold_id, 5, y_class=y.astype(np.int32))
cands = W.candidates(w)
r = eng.svc(["rbf"] * len(cands), [c["C"] for c in cands], np.array([c["gamma"] for c in cands])[:, None])
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", w["name"] + ".npz"), allow_pickle=True)
gi = g["diag"][:, :, 0].astype(np.int64)
bad = np.argwhere(r["n_iter"] != gi)
gam = sorted(set(float(c["gamma"]) for c in cands)); Cs = sorted(set(float(c["C"]) for c in cands))
print("mismatches", len(bad), "of", gi.size, "| score mismatches", (r["test"] != g["test_scores"]).sum(), (r["train"] != g["train_scores"]).sum())
for c, f in bad:
    print("cand %3d C#%2d gamma#%2d fold %d: gpu %6d sklearn %6d" % (c, Cs.index(float(cands[c]["C"])), gam.index(float(cands[c]["gamma"])), f, r["n_iter"][c, f], gi[c, f]))
print("profile", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in eng.profile().items() if k.startswith("ms_")})
