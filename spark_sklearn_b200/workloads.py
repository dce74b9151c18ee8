"""Synthetic workloads for the BASELINE.json configs (SURVEY.md §8d recipes).

Every recipe is deterministic (``random_state=0``), fp32, C-contiguous.  They are
shared by ``bench.py``, the parity tests and ``tests/golden/make_goldens.py`` so
that the GPU path, the oracle and the committed goldens all see identical bytes.

A workload is a dict: ``X, y, estimator (name), param_grid | param_distributions,
cv, search ("grid" | "random"), n_iter, random_state``.
"""
import numpy as np

__all__ = ["make_workload", "WORKLOADS"]


def _c1():
    from sklearn.datasets import load_iris
    iris = load_iris()
    return dict(name="c1_iris_svc", X=np.ascontiguousarray(iris.data), y=iris.target,
                estimator="SVC", est_params={"gamma": "auto"},
                param_grid={"kernel": ("linear", "rbf"), "C": [1, 10]}, cv=5, search="grid")


def _svc_data(n=10000, d=512):
    from sklearn.datasets import make_classification
    from sklearn.preprocessing import StandardScaler
    X, y = make_classification(n_samples=n, n_features=d, n_informative=d // 8, n_redundant=0,
                               n_classes=2, class_sep=1.0, flip_y=0.01, random_state=0)
    X = StandardScaler().fit_transform(X).astype(np.float32)
    return np.ascontiguousarray(X), y.astype(np.int64)


def _c2(n=10000, d=512, nc=8, ng=8, cv=5, name="c2_svc_rbf_8x8"):
    X, y = _svc_data(n, d)
    grid = {"C": np.logspace(-1, 2.5, nc), "gamma": np.geomspace(8.0 / (d * 64), 8.0 / (d * 4), ng)}
    # d=512 -> gamma in [1/4096, 1/256] exactly as SURVEY.md §8d states
    return dict(name=name, X=X, y=y, estimator="SVC", est_params={"kernel": "rbf"},
                param_grid=grid, cv=cv, search="grid")


def _c4():
    return _c2(nc=16, ng=16, name="c4_svc_rbf_16x16")


def _c3(n=50000, d=256, n_iter=256, cv=5, name="c3_logreg_random256"):
    from sklearn.datasets import make_classification
    from sklearn.preprocessing import StandardScaler
    from scipy.stats import loguniform
    X, y = make_classification(n_samples=n, n_features=d, n_informative=d // 8, n_redundant=0,
                               n_classes=2, class_sep=0.5, flip_y=0.02, random_state=0)
    X = np.ascontiguousarray(StandardScaler().fit_transform(X).astype(np.float32))
    return dict(name=name, X=X, y=y.astype(np.int64), estimator="LogisticRegression", est_params={},
                param_distributions={"C": loguniform(1e-4, 1e2)}, n_iter=n_iter, random_state=0,
                cv=cv, search="random")


def _c5(n=20000, d=1024, n_alpha=512, cv=10, name="c5_ridge_512"):
    from sklearn.datasets import make_regression
    X, y = make_regression(n_samples=n, n_features=d, n_informative=d // 8, noise=10.0, random_state=0)
    return dict(name=name, X=np.ascontiguousarray(X.astype(np.float32)), y=y.astype(np.float32),
                estimator="Ridge", est_params={}, param_grid={"alpha": np.logspace(-3, 5, n_alpha)},
                cv=cv, search="grid")


def _lasso(n=20000, d=1024, n_alpha=32, cv=10, name="lasso_1024", enet=False):
    """The estimator of the reference's own search tests (tests/test_search_2.py:69-119) on config 5's data recipe."""
    w = _c5(n=n, d=d, n_alpha=2, cv=cv, name=name)
    if enet:
        w.update(estimator="ElasticNet", param_grid={"alpha": np.logspace(-2, 1.5, n_alpha), "l1_ratio": [0.2, 0.7]})
    else:
        w.update(estimator="Lasso", param_grid={"alpha": np.logspace(-2, 2, n_alpha)})
    return w


WORKLOADS = {
    "c1": _c1, "c2": _c2, "c3": _c3, "c4": _c4, "c5": _c5,
    # reduced-size variants: same recipes, sizes the CPU oracle finishes in seconds
    "c2_small": lambda: _c2(n=1000, d=64, nc=4, ng=4, name="c2_small"),
    "c2_mid": lambda: _c2(n=3000, d=128, nc=4, ng=4, name="c2_mid"),
    "c3_small": lambda: _c3(n=4000, d=32, n_iter=16, name="c3_small"),
    "c5_small": lambda: _c5(n=2000, d=64, n_alpha=32, cv=10, name="c5_small"),
    # Lasso / ElasticNet on the fold Grams
    "lasso_1024": _lasso,
    "lasso_small": lambda: _lasso(n=2000, d=64, n_alpha=16, cv=5, name="lasso_small"),
    "enet_small": lambda: _lasso(n=1500, d=200, n_alpha=6, cv=4, name="enet_small", enet=True),
}


def make_workload(key):
    return WORKLOADS[key]()


def make_estimator(w):
    """Instantiate the sklearn estimator a workload names."""
    if w["estimator"] == "SVC":
        from sklearn.svm import SVC
        return SVC(**w["est_params"])
    if w["estimator"] == "LogisticRegression":
        from sklearn.linear_model import LogisticRegression
        return LogisticRegression(**w["est_params"])
    if w["estimator"] == "Ridge":
        from sklearn.linear_model import Ridge
        return Ridge(**w["est_params"])
    if w["estimator"] in ("Lasso", "ElasticNet"):
        import sklearn.linear_model as lm
        return getattr(lm, w["estimator"])(**w["est_params"])
    raise ValueError(w["estimator"])


def candidates(w):
    """The candidate list exactly as the reference enumerates it
    (ParameterGrid: reference grid_search.py:246; ParameterSampler: random_search.py:222)."""
    from sklearn.model_selection import ParameterGrid, ParameterSampler
    if w["search"] == "grid":
        return list(ParameterGrid(w["param_grid"]))
    return list(ParameterSampler(w["param_distributions"], w["n_iter"], random_state=w["random_state"]))
