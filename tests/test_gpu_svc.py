"""Kernel C-SVC (one-vs-one) against libsvm itself (sklearn.svm.SVC, the class the reference
scripts instantiate: scripts/aligned_decode_svm_ncv.py:313-317, aligned_decode_svm.py:262-263)
on seeded inputs, through the C ABI."""
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(seed, n, k, ncls, sep=1.0, weights=None):
    rng = np.random.default_rng(seed)
    y = rng.choice(ncls, size=n, p=weights) + 1
    cent = rng.standard_normal((ncls + 1, k)) * sep
    X = cent[y] + rng.standard_normal((n, k))
    return X.astype(np.float32).astype(np.float64), y


CASES = [
    dict(n=300, k=20, ncls=5, kernel='rbf', class_weight='balanced', sep=1.0),
    dict(n=1145, k=65, ncls=9, kernel='rbf', class_weight='balanced', sep=0.6,
         weights=[.2, .05, .1, .15, .1, .1, .1, .12, .08]),
    dict(n=260, k=60, ncls=9, kernel='rbf', class_weight=None, sep=0.5),
    dict(n=200, k=10, ncls=3, kernel='linear', class_weight=None, sep=0.8),
    dict(n=400, k=12, ncls=2, kernel='rbf', class_weight='balanced', sep=0.7, weights=[.8, .2]),
    dict(n=150, k=8, ncls=2, kernel='linear', class_weight=None, sep=0.5),
]


@pytest.mark.parametrize('case', CASES)
def test_svc_matches_libsvm(lib_built, case):
    from sklearn.svm import SVC as RefSVC
    from cross_patient_speech_decoding_b200.svm import SVC
    X, y = _data(3, case['n'], case['k'], case['ncls'], case['sep'], case.get('weights'))
    Xte, yte = _data(4, 200, case['k'], case['ncls'], case['sep'], case.get('weights'))
    # same stopping tolerance on both sides; tight so the optimum (unique) is what is compared
    kw = dict(kernel=case['kernel'], class_weight=case['class_weight'], tol=1e-6)
    ref = RefSVC(decision_function_shape='ovo', **kw).fit(X, y)
    got = SVC(decision_function_shape='ovo', **kw).fit(X, y)
    assert np.array_equal(got.classes_, ref.classes_)
    if case['kernel'] == 'rbf':
        assert abs(got._gamma - ref._gamma) <= 1e-9 * ref._gamma
    assert np.abs(got.intercept_ - ref.intercept_).max() < 2e-4
    d_ref = ref.decision_function(Xte)
    d_got = got.decision_function(Xte)
    assert np.abs(d_got - d_ref).max() < 1e-3 * max(1.0, np.abs(d_ref).max())
    assert np.mean(got.predict(Xte) == ref.predict(Xte)) >= 0.995
    assert np.array_equal(got.predict(X), ref.predict(X)) or np.mean(got.predict(X) == ref.predict(X)) >= 0.995
    # support vectors and dual coefficients in libsvm's layout
    assert np.array_equal(got.n_support_, ref.n_support_) or \
        np.abs(got.n_support_ - ref.n_support_).max() <= 2
    if np.array_equal(got.support_, ref.support_):
        assert np.abs(got.dual_coef_ - ref.dual_coef_).max() < 5e-3


def test_svc_default_tolerance_labels(lib_built):
    """libsvm's default tol=1e-3 on both sides: the two runs stop at slightly different points of
    the same dual, labels agree on >= 99 % (the north-star bar)."""
    from sklearn.svm import SVC as RefSVC
    from cross_patient_speech_decoding_b200.svm import SVC
    X, y = _data(7, 1100, 65, 9, 0.6)
    Xte, _ = _data(8, 400, 65, 9, 0.6)
    ref = RefSVC(kernel='rbf', class_weight='balanced').fit(X, y)
    got = SVC(kernel='rbf', class_weight='balanced').fit(X, y)
    assert np.mean(got.predict(Xte) == ref.predict(Xte)) >= 0.99
    assert got.fit_status_ == 0
    # sklearn's one-vs-rest shaped decision function (votes + squashed confidences)
    assert np.abs(got.decision_function(Xte) - ref.decision_function(Xte)).max() < 5e-3


def test_svc_sklearn_plumbing(lib_built):
    """clone / get_params / pickle, and sklearn's own BaggingClassifier driving the class (the
    reference's BaggingClassifier(SVC(kernel='linear'), 10), aligned_decode_svm.py:262-263)."""
    from sklearn.base import clone
    from sklearn.ensemble import BaggingClassifier
    from sklearn.svm import SVC as RefSVC
    from cross_patient_speech_decoding_b200.svm import SVC
    X, y = _data(11, 240, 12, 4, 1.2)
    Xte, _ = _data(12, 120, 12, 4, 1.2)
    m = SVC(kernel='linear')
    assert clone(m).get_params()['kernel'] == 'linear'
    m.fit(X, y)
    m2 = pickle.loads(pickle.dumps(m))
    assert np.array_equal(m2.predict(Xte), m.predict(Xte))
    bag = BaggingClassifier(estimator=SVC(kernel='linear'), n_estimators=10, random_state=5).fit(X, y)
    ref = BaggingClassifier(estimator=RefSVC(kernel='linear'), n_estimators=10, random_state=5).fit(X, y)
    # same bootstrap streams (sklearn draws them); members agree up to the tol=1e-3 stopping point
    for a, b in zip(bag.estimators_samples_, ref.estimators_samples_):
        assert np.array_equal(a, b)
    assert np.mean(bag.predict(Xte) == ref.predict(Xte)) >= 0.97


def test_svc_errors(lib_built):
    from cross_patient_speech_decoding_b200.svm import SVC
    X, y = _data(1, 40, 4, 3)
    with pytest.raises(ValueError):
        SVC().fit(X, np.ones(40, dtype=int))
    with pytest.raises(ValueError):
        SVC(kernel='poly').fit(X, y)
    with pytest.raises(ValueError):
        SVC(class_weight={1: 2.0}).fit(X, y)
