"""CPU: pin the oracle (oracle/) against scikit-learn itself and against the committed goldens.

The reference's own tests pin no numeric result on this path; its arithmetic is
scikit-learn's, so the pin is scikit-learn 1.9.0 run in this image plus tests/golden/*.npz made by
tests/golden/make_goldens.py.
"""
import numpy as np
import pytest
from sklearn.model_selection import StratifiedKFold
from sklearn.svm import SVC

from conftest import golden
from oracle import oracle as O
from spark_sklearn_b200 import workloads as W


def _compare_with_sklearn(X, y, train, test, exact=True, **prm):
    X64 = np.ascontiguousarray(X, np.float64)
    m = O.SVCModel(X64, y, train, **prm)
    s = SVC(**prm).fit(X[train], y[train])
    assert list(s.n_iter_) == m.n_iter                      # same iterate sequence, not just same optimum
    rho = np.array([p[4] for p in m.pairs])
    if exact:
        np.testing.assert_array_equal(rho, -s._intercept_)
    else:       # float64 features: products are not exact, BLAS ddot's summation order shows at 1e-15
        np.testing.assert_allclose(rho, -s._intercept_, rtol=1e-12)
    if len(m.classes) == 2:
        _, _, rows, coef, _ = m.pairs[0]
        full = np.zeros(len(X)); full[rows] = coef
        sk = np.zeros(len(X)); sk[np.asarray(train)[s.support_]] = s._dual_coef_[0]
        np.testing.assert_array_equal(full, sk)             # bit-exact dual coefficients
    np.testing.assert_array_equal(m.predict(test), s.predict(X[test]))
    np.testing.assert_array_equal(m.predict(train), s.predict(X[train]))


def test_svc_oracle_bitexact_iris():
    w = W.make_workload("c1")
    X, y = w["X"], w["y"]
    tr, te = next(iter(StratifiedKFold(5).split(X, y)))
    for kernel in ("linear", "rbf"):
        for C in (1, 10):
            _compare_with_sklearn(X, y, tr, te, exact=False, kernel=kernel, C=C, gamma="auto")


@pytest.mark.parametrize("C,gamma", [(0.1, 1 / 64), (10.0, 1 / 128), (316.0, 1 / 1024), (1.0, "scale")])
def test_svc_oracle_bitexact_with_shrinking(C, gamma):
    w = W.make_workload("c2_mid")                          # l = 2400 > 1000: shrinking and unshrinking happen
    X, y = w["X"], w["y"]
    tr, te = next(iter(StratifiedKFold(5).split(X, y)))
    _compare_with_sklearn(X, y, tr, te, kernel="rbf", C=C, gamma=gamma)


def test_svc_oracle_vs_golden_c1_and_survey_table():
    g = golden("c1_iris_svc")
    # SURVEY.md §8c / BASELINE.md §2 table (generated during the survey with sklearn GridSearchCV)
    np.testing.assert_allclose(g["mean_test_score"], [0.98, 0.98, 0.9733333333, 0.98], atol=1e-9)
    np.testing.assert_allclose(g["mean_train_score"], [0.9816666667, 0.9833333333, 0.9783333333, 0.9766666667], atol=1e-9)
    w = W.make_workload("c1")
    fold_id, ns = O.folds_from_cv(w["cv"], w["X"], w["y"], True)
    np.testing.assert_array_equal(fold_id, g["fold_id"])
    test, train, _ = O.cv_scores_svc(w["X"], w["y"], fold_id, ns, W.candidates(w), w["est_params"])
    np.testing.assert_array_equal(test, g["test_scores"])
    np.testing.assert_array_equal(train, g["train_scores"])


def test_svc_oracle_vs_golden_c2_small():
    g = golden("c2_small")
    w = W.make_workload("c2_small")
    cands = W.candidates(w)[::3]
    fold_id, ns = O.folds_from_cv(w["cv"], w["X"], w["y"], True)
    test, train, iters = O.cv_scores_svc(w["X"], w["y"], fold_id, ns, cands, w["est_params"])
    np.testing.assert_array_equal(test, g["test_scores"][::3])
    np.testing.assert_array_equal(train, g["train_scores"][::3])
    np.testing.assert_array_equal(iters, g["diag"][::3, :, 0].astype(np.int64))


def test_golden_tasks_mode_equals_sklearn_search():
    """make_goldens 'tasks' mode (the reference's task list) == sklearn GridSearchCV run as a whole."""
    from sklearn.model_selection import GridSearchCV
    w = W.make_workload("c2_small")
    g = golden("c2_small")
    s = GridSearchCV(W.make_estimator(w), w["param_grid"], cv=w["cv"], return_train_score=True).fit(w["X"], w["y"])
    for k in range(5):
        np.testing.assert_array_equal(s.cv_results_["split%d_test_score" % k], g["test_scores"][:, k])
        np.testing.assert_array_equal(s.cv_results_["split%d_train_score" % k], g["train_scores"][:, k])


def test_ridge_oracle_vs_golden():
    g = golden("c5_small")
    w = W.make_workload("c5_small")
    fold_id, ns = O.folds_from_cv(w["cv"], w["X"], w["y"], False)
    test, train = O.cv_scores_ridge(w["X"], w["y"], fold_id, ns, W.candidates(w)[::4])
    np.testing.assert_allclose(test, g["test_scores"][::4], atol=2e-6)
    np.testing.assert_allclose(train, g["train_scores"][::4], atol=2e-6)


def test_logreg_oracle_vs_golden():
    g = golden("c3_small")
    w = W.make_workload("c3_small")
    fold_id, ns = O.folds_from_cv(w["cv"], w["X"], w["y"], True)
    test, train = O.cv_scores_logreg(w["X"], w["y"], fold_id, ns, W.candidates(w)[::4])
    # L-BFGS stops early (gtol); the restated objective follows the same path up to float32 rounding
    assert np.abs(test - g["test_scores"][::4]).max() <= 2.5e-3
    assert np.abs(test.mean(1) - g["test_scores"][::4].mean(1)).max() <= 1e-3


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_enet_oracle_matches_sklearn_iterates(dtype):
    """The coordinate-descent restatement takes scikit-learn's own sweeps: same n_iter_, same coefficients (float64:
    to the last bit of the score; float32: BLAS summation order shows at 1e-7)."""
    import warnings
    from sklearn.datasets import make_regression
    from sklearn.linear_model import ElasticNet, Lasso
    X, y = make_regression(n_samples=400, n_features=60, n_informative=10, noise=5.0, random_state=0)
    X, y = X.astype(dtype), y.astype(dtype)
    tr, te = np.arange(300), np.arange(300, 400)
    for alpha, l1 in ((0.01, 1.0), (1.0, 1.0), (30.0, 1.0), (0.5, 0.5), (0.1, 0.0), (200.0, 1.0)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = (Lasso(alpha=alpha) if l1 == 1.0 else ElasticNet(alpha=alpha, l1_ratio=l1)).fit(X[tr], y[tr])
        t, r, it = O.enet_fit_score(X, y, tr, te, alpha, l1)
        assert it == m.n_iter_, (alpha, l1)
        tol = 0 if dtype is np.float64 else 2e-7
        assert abs(m.score(X[te], y[te]) - t) <= tol and abs(m.score(X[tr], y[tr]) - r) <= tol, (alpha, l1)


@pytest.mark.parametrize("key", ["lasso_small", "enet_small"])
def test_enet_oracle_vs_golden(key):
    g = golden(key)
    w = W.make_workload(key)
    fold_id, ns = O.folds_from_cv(w["cv"], w["X"], w["y"], False)
    cands = W.candidates(w)[::3]
    test, train, iters = O.cv_scores_enet(w["X"], w["y"], fold_id, ns, cands)
    np.testing.assert_allclose(test, g["test_scores"][::3], atol=2e-6)
    np.testing.assert_allclose(train, g["train_scores"][::3], atol=2e-6)
    np.testing.assert_array_equal(iters, g["diag"][::3, :, 0].astype(int))


def test_multinomial_logreg_oracle_vs_sklearn():
    """Three and more classes: the restatement follows scikit-learn's float32 arithmetic closely enough for the same
    scores; L-BFGS-B stops after a few dozen iterations, where float32 rounding can move the stop by one iteration."""
    import warnings
    from sklearn.datasets import load_iris, make_classification
    from sklearn.linear_model import LogisticRegression
    same = total = 0
    for seed in range(2):
        X, y = make_classification(n_samples=1500, n_features=20, n_informative=10, n_classes=4, n_clusters_per_class=1, random_state=seed)
        X = X.astype(np.float32)
        tr, te = np.arange(1000), np.arange(1000, 1500)
        for C in (0.01, 1.0, 100.0):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                m = LogisticRegression(C=C).fit(X[tr], y[tr])
            t, r, it = O.logreg_fit_score(X, y, tr, te, C)
            assert abs(it - m.n_iter_[0]) <= 1
            same += it == m.n_iter_[0]; total += 1
            assert abs(m.score(X[te], y[te]) - t) <= 2.1 / len(te) and abs(m.score(X[tr], y[tr]) - r) <= 2.1 / len(tr)
    assert same >= total - 2
    Xi, yi = load_iris(return_X_y=True)                     # float64: identical iterates
    idx = np.random.RandomState(0).permutation(150)
    tr, te = idx[:100], idx[100:]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = LogisticRegression(C=0.1).fit(Xi[tr], yi[tr])
    t, r, it = O.logreg_fit_score(Xi, yi, tr, te, 0.1)
    assert it == m.n_iter_[0] and t == m.score(Xi[te], yi[te]) and r == m.score(Xi[tr], yi[tr])
