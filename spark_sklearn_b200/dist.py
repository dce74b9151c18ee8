"""Multi-GPU plumbing: one process per GPU (torchrun), candidates strided over ranks, ONE all-gather
of the per-candidate score blocks -- the counterpart of ``RDD.collect()`` (reference
base_search.py:89).  The data path has no other collective: tasks are independent (reference
base_search.py:56-62 runs one Spark partition per task) and the dataset is replicated on every
GPU exactly as the reference replicates it with ``sc.broadcast`` (base_search.py:63-65).
"""
import numpy as np


def _td():
    try:
        import torch.distributed as td
    except Exception:  # pragma: no cover
        return None
    return td if td.is_available() and td.is_initialized() else None


def rank_world():
    td = _td()
    return (td.get_rank(), td.get_world_size()) if td else (0, 1)


def allgather_candidates(local, my, n_cand, n_splits, world):
    """local: dict of [len(my), n_splits] arrays (test, train|None, fit_time, score_time) for the
    candidates ``my`` of this rank.  Returns the same dict for all n_cand candidates, identical on
    every rank and independent of the number of ranks."""
    keys = ["test", "train", "fit_time", "score_time"]
    if world == 1:
        return {k: local.get(k) for k in keys}
    import torch
    td = _td()
    per = (n_cand + world - 1) // world                      # pad to equal counts
    buf = np.full((per, n_splits, len(keys)), np.nan)
    for j, k in enumerate(keys):
        if local.get(k) is not None and len(my):
            buf[:len(my), :, j] = local[k]
    dev = torch.device("cuda", torch.cuda.current_device()) if td.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(buf).to(dev)
    gathered = torch.empty((world * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
    td.all_gather_into_tensor(gathered, t.contiguous())
    g = gathered.cpu().numpy().reshape((world, per) + tuple(t.shape[1:]))
    out = {}
    for j, k in enumerate(keys):
        if local.get(k) is None and k == "train":
            out[k] = None
            continue
        full = np.empty((n_cand, n_splits))
        for r in range(world):
            idx = list(range(r, n_cand, world))
            full[idx] = g[r, :len(idx), :, j]
        out[k] = full
    return out
