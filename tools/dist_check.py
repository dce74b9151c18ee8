"""torchrun --nproc-per-node N tools/dist_check.py : the same search on N GPUs must give the golden cv_results_ on every rank."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spark_sklearn_b200 import GridSearchCV, workloads as W
lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
w = W.make_workload("c2_mid")
s = GridSearchCV(None, W.make_estimator(w), w["param_grid"], cv=w["cv"]).fit(w["X"], w["y"])
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c2_mid.npz"))
got = np.stack([s.cv_results_["split%d_test_score" % k] for k in range(5)], 1)
ok = np.array_equal(got, g["test_scores"]) and s.best_index_ == int(np.flatnonzero(s.cv_results_["rank_test_score"] == 1)[0])
pred = s.predict(w["X"][:200])
print("rank %d/%d: scores == golden: %s, best_index %d, predict ok %s" % (dist.get_rank(), dist.get_world_size(), ok, s.best_index_, pred.shape), flush=True)
assert ok
dist.destroy_process_group()
