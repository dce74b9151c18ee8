"""GPU parity tests for the LogisticRegression path (batched L-BFGS-B restatement on tensor-core GEMMs).

Checker: goldens made by scikit-learn 1.9.0 (scipy L-BFGS-B on float32 loss/gradient).  The optimiser stops early
(gtol=1e-4), so parity needs the same trajectory: iteration counts must match scipy's on (almost) every fit and
scores must agree to a flip or two (BASELINE asks 1e-4 on mean_test_score at 50000 rows; one flip on a fold of the
reduced 4000-row workload is 1.25e-3)."""
import numpy as np
import pytest

from conftest import golden
from spark_sklearn_b200 import workloads as W

pytestmark = pytest.mark.gpu


def test_logreg_c3_small_vs_golden(engine):
    from oracle import oracle as O
    w = W.make_workload("c3_small")
    fold_id, ns = O.folds_from_cv(w["cv"], w["X"], w["y"], True)
    engine.set_data(w["X"], fold_id, ns, y_class=w["y"].astype(np.int32))
    g = golden("c3_small")
    r = engine.logreg([c["C"] for c in W.candidates(w)])
    it_gold = g["diag"][:, :, 0].astype(np.int32)
    assert np.mean(r["n_iter"] == it_gold) >= 0.9, (r["n_iter"], it_gold)      # same L-BFGS-B trajectory
    assert np.abs(r["test"] - g["test_scores"]).max() <= 2 * 1.25e-3 + 1e-12       # at most two flips on an 800-row fold
    assert np.abs(r["test"].mean(1) - g["test_scores"].mean(1)).max() <= 5e-4 + 1e-12
    assert np.abs(r["train"] - g["train_scores"]).max() <= 4 * 3.2e-4 + 1e-12      # <= 4 flips on 3200 training rows


def test_logreg_python_api_and_refit(engine):
    """Public API + refit.  The optimiser is trajectory-identical to scipy's (coefficients agree to ~1e-6) except when
    the gtol stopping test is borderline: |g|_inf within float32 rounding of 1e-4 stops one iteration earlier or later
    than scikit-learn's own float32 BLAS run does (measured: 1 of 40 reduced-size fits), worth a flip or two."""
    from sklearn.linear_model import LogisticRegression
    from sklearn.model_selection import GridSearchCV as SkGrid
    from spark_sklearn_b200 import GridSearchCV
    w = W.make_workload("c3_small")
    X, y = w["X"], w["y"]
    grid = {"C": [1e-3, 1e-1, 50.0]}
    a = GridSearchCV(None, LogisticRegression(), grid, cv=5).fit(X, y)
    b = SkGrid(LogisticRegression(), grid, cv=5, return_train_score=True).fit(X, y)
    assert np.abs(a.cv_results_["mean_test_score"] - b.cv_results_["mean_test_score"]).max() <= 5e-4 + 1e-12   # <= 2 flips / 4000
    ca, cb = a.best_estimator_.coef_, b.best_estimator_.coef_
    assert a.best_params_ == b.best_params_
    assert a.best_estimator_.n_iter_[0] == b.best_estimator_.n_iter_[0]
    assert np.abs(ca - cb).max() <= 1e-5 * np.abs(cb).max()
    assert abs(a.best_estimator_.intercept_[0] - b.best_estimator_.intercept_[0]) <= 1e-6
    np.testing.assert_array_equal(a.predict(X), b.predict(X))
    assert a.predict_proba(X).shape == (len(y), 2)


def test_logreg_c3_full_size_vs_golden(engine):
    """BASELINE config 3 at full size: 256 candidates x 5 folds on 50000x256 -- the 1e-4 bar on mean_test_score."""
    from oracle import oracle as O
    w = W.make_workload("c3")
    fold_id, ns = O.folds_from_cv(w["cv"], w["X"], w["y"], True)
    engine.set_data(w["X"], fold_id, ns, y_class=w["y"].astype(np.int32))
    g = golden("c3_logreg_random256")
    r = engine.logreg([c["C"] for c in W.candidates(w)])
    assert np.mean(r["n_iter"] == g["diag"][:, :, 0].astype(np.int32)) >= 0.995
    assert np.abs(r["test"].mean(1) - g["test_scores"].mean(1)).max() <= 1e-4 + 1e-12


def test_logreg_class_weight_vs_sklearn(engine):
    """LogisticRegression(class_weight=...): sample_weight = class_weight_[y] multiplies the pointwise loss and gradient and
    replaces n by the weight sum in the l2 scaling (_logistic.py); 'balanced' is recomputed from every fold's training labels."""
    from sklearn.linear_model import LogisticRegression
    from sklearn.model_selection import GridSearchCV as SkGrid
    from spark_sklearn_b200 import GridSearchCV
    w = W.make_workload("c3_small")
    X, y = w["X"], w["y"].copy()
    y[:900] = 0                                                     # unbalanced
    grid = {"C": [1e-2, 1.0], "class_weight": [None, "balanced", {0: 1.0, 1: 3.0}]}
    a = GridSearchCV(None, LogisticRegression(), grid, cv=5).fit(X, y)
    b = SkGrid(LogisticRegression(), grid, cv=5, return_train_score=True).fit(X, y)
    assert np.abs(a.cv_results_["mean_test_score"] - b.cv_results_["mean_test_score"]).max() <= 7.5e-4      # <= 3 flips / 4000
    assert np.abs(a.cv_results_["mean_train_score"] - b.cv_results_["mean_train_score"]).max() <= 7.5e-4
    assert a.best_params_ == b.best_params_
    ca, cb = a.best_estimator_.coef_, b.best_estimator_.coef_
    assert np.abs(ca - cb).max() <= 2e-4 * np.abs(cb).max()
    assert np.mean(a.predict(X) != b.predict(X)) <= 1e-3


def _multiclass_data(n=3000, d=24, k=4, seed=0):
    from sklearn.datasets import make_classification
    X, y = make_classification(n_samples=n, n_features=d, n_informative=12, n_classes=k, n_clusters_per_class=1,
                               class_sep=0.8, flip_y=0.02, random_state=seed)
    return X.astype(np.float32), y


def test_multinomial_logreg_vs_oracle_and_sklearn(engine):
    """Three and more classes (scikit-learn: multinomial loss, one weight row per class): the same batched L-BFGS-B with
    n_classes weight rows per fit.  Checker: scikit-learn itself and the float32-faithful restatement.  Iteration counts
    within one or two of scipy's (float32 rounding moves the early stop), scores within a few flips."""
    import warnings
    from oracle import oracle as O
    from sklearn.linear_model import LogisticRegression
    X, y = _multiclass_data()
    fold_id, ns = O.folds_from_cv(4, X, y, True)
    engine.set_data(X, fold_id, ns, y_class=y.astype(np.int32))
    Cs = [1e-3, 0.05, 1.0, 30.0]
    r = engine.logreg(Cs)
    te, tr, it = O.cv_scores_logreg(X, y, fold_id, ns, [{"C": c} for c in Cs], return_n_iter=True)
    n_te, n_tr = len(y) // ns, len(y) - len(y) // ns
    # the loss sum is accumulated with floating-point atomics (run-to-run order), so the early gtol stop may move by an
    # iteration between runs as well as against scipy: a handful of flips either way
    assert np.abs(r["test"] - te).max() <= 4.0 / n_te + 1e-12, (r["test"], te)
    assert np.abs(r["train"] - tr).max() <= 8.0 / n_tr + 1e-12, (r["train"], tr)
    assert np.abs(r["test"].mean(1) - te.mean(1)).max() <= 1.5e-3
    assert np.abs(r["n_iter"] - it).max() <= 4 and np.mean(np.abs(r["n_iter"] - it) <= 1) >= 0.7, (r["n_iter"], it)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for ci, c in enumerate(Cs[1:3], 1):
            for k in range(ns):
                m = LogisticRegression(C=c).fit(X[fold_id != k], y[fold_id != k])
                assert abs(m.score(X[fold_id == k], y[fold_id == k]) - r["test"][ci, k]) <= 4.0 / n_te + 1e-12
                assert abs(int(m.n_iter_[0]) - int(r["n_iter"][ci, k])) <= 4


def test_multinomial_logreg_python_api_iris_and_scorers(engine):
    from sklearn.datasets import load_iris
    from sklearn.linear_model import LogisticRegression
    from sklearn.model_selection import GridSearchCV as SkGrid
    from spark_sklearn_b200 import GridSearchCV
    import warnings
    Xi, yi = load_iris(return_X_y=True)
    grid = {"C": [0.1, 1.0, 10.0]}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = GridSearchCV(None, LogisticRegression(max_iter=200), grid, cv=5).fit(Xi, yi)
        b = SkGrid(LogisticRegression(max_iter=200), grid, cv=5, return_train_score=True).fit(Xi, yi)
    assert np.abs(a.cv_results_["mean_test_score"] - b.cv_results_["mean_test_score"]).max() <= 1.0 / 30 / 5 * 3 + 1e-12   # <= 3 flips over the 5 folds
    ea, eb = a.best_estimator_, b.best_estimator_
    assert ea.coef_.shape == eb.coef_.shape == (3, 4) and ea.intercept_.shape == (3,)
    assert (a.predict(Xi) == eb.predict(Xi)).mean() >= 0.98
    np.testing.assert_allclose(ea.predict_proba(Xi), eb.predict_proba(Xi), atol=0.05)
    X, y = _multiclass_data(n=2000, seed=1)
    for scoring in ("f1_macro", "balanced_accuracy", "f1_weighted"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a = GridSearchCV(None, LogisticRegression(class_weight="balanced"), {"C": [0.01, 1.0]}, cv=4, scoring=scoring).fit(X, y)
            b = SkGrid(LogisticRegression(class_weight="balanced"), {"C": [0.01, 1.0]}, cv=4, scoring=scoring, return_train_score=True).fit(X, y)
        assert np.abs(a.cv_results_["mean_test_score"] - b.cv_results_["mean_test_score"]).max() <= 4e-3, scoring
        assert np.abs(a.cv_results_["mean_train_score"] - b.cv_results_["mean_train_score"]).max() <= 4e-3, scoring
    with pytest.raises(NotImplementedError):
        GridSearchCV(None, LogisticRegression(), {"C": [1.0]}, cv=3, scoring="roc_auc").fit(X, y)


def test_logreg_sample_weight_fit_params(engine):
    """fit_params={'sample_weight': w}: weighted loss and gradient, weight-sum scaling of the penalty, unweighted scores (the
    reference's task: fit_params go to fit() only); with class_weight='balanced' the class frequencies are counted by
    weight.  Binary and multinomial."""
    import warnings
    from sklearn.linear_model import LogisticRegression
    from spark_sklearn_b200 import GridSearchCV
    from test_gpu_enet import _reference_tasks
    w3 = W.make_workload("c3_small")
    Xm, ym = _multiclass_data(n=2400, seed=2)
    for X, y, est in ((w3["X"], w3["y"], LogisticRegression()), (Xm, ym, LogisticRegression()),
                      (w3["X"], w3["y"], LogisticRegression(class_weight="balanced"))):
        rng = np.random.RandomState(7)
        sw = rng.gamma(1.0, 1.0, len(y)) + 0.05
        grid = {"C": [1e-2, 1.0]}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a = GridSearchCV(None, est, grid, cv=4, iid=False, fit_params={"sample_weight": sw}).fit(X, y)
            te, tr, best, fitted = _reference_tasks(est, grid, 4, X, y, sw, classifier=True)
        n_te = len(y) // 4
        assert np.abs(a.cv_results_["mean_test_score"] - te).max() <= 2.0 / n_te
        assert np.abs(a.cv_results_["mean_train_score"] - tr).max() <= 2.0 / n_te
        if a.best_params_ == best:
            np.testing.assert_allclose(a.best_estimator_.coef_, fitted.coef_, atol=0.02 * np.abs(fitted.coef_).max() + 1e-3)
