"""CPU checks of the test oracle itself (no GPU): the restated MCCA is pinned by known
answers against importable reference code, the CPU port of the path reproduces the golden
outputs the real reference produced, and the exact SVM solver agrees with liblinear."""
import os
import sys
import warnings

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))

from oracle import mcca_restated, pipeline_port, reference_path, svm_exact  # noqa: E402


def _views(seed, n=400, dims=(7, 5, 6)):
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((n, 4))
    return [Z @ rng.standard_normal((4, d)) + 0.4 * rng.standard_normal((n, d)) for d in dims]


def test_mcca_two_views_unregularised_equals_cca():
    """Known answer: for 2 views, regs=None, full rank, the GEVP eigenvalues are 1 + rho with
    rho the canonical correlations (of the reference's CCA_align when it is importable)."""
    Xa, Xb = _views(0, dims=(6, 6))
    m = mcca_restated.MCCARestated(n_components=6, regs=None).fit([Xa, Xb])
    if reference_path.available():
        ref = reference_path.load()
        _, _, rho = ref.CCA_align(Xa.T.copy(), Xb.T.copy())
    else:
        _, _, rho = pipeline_port.cca_directions(Xa, Xb)
    assert np.abs((m.evals_ - 1.0) - rho).max() < 1e-10
    # canonical variates of the two views span the same subspaces as CCA's
    Ma, Mb, _ = pipeline_port.cca_directions(Xa, Xb)
    Ya, Za = (Xa - Xa.mean(0)) @ Ma, m.transform_view(Xa, 0)
    s = np.linalg.svd(np.linalg.qr(Ya)[0].T @ np.linalg.qr(Za)[0], compute_uv=False)
    assert s.min() > 1 - 1e-10


def test_mcca_gevp_residual_and_normalisation():
    views = _views(1)
    Xc = [v - v.mean(0) for v in views]
    for regs in (None, 0.5):
        m = mcca_restated.MCCARestated(n_components=5, regs=regs).fit(views)
        lhs, rhs, _ = mcca_restated._gevp_blocks(Xc, regs)
        W = np.vstack(m.loadings_)
        assert np.abs(lhs @ W - (rhs @ W) * m.evals_).max() <= 1e-8 * np.abs(lhs).max()
        assert np.abs(W.T @ rhs @ W - np.eye(5)).max() < 1e-8
        assert np.all(np.diff(m.evals_) <= 1e-12)


def test_mcca_informative_equals_gevp_on_reduced_views():
    """signal_ranks path: loadings map back through the per-view SVD bases, and full ranks
    reproduce the plain GEVP solution (same eigenvalues)."""
    views = _views(2)
    full = mcca_restated.MCCARestated(n_components=4, regs=0.5).fit(views)
    red = mcca_restated.MCCARestated(n_components=4, regs=0.5,
                                     signal_ranks=[v.shape[1] for v in views]).fit(views)
    assert np.abs(full.evals_ - red.evals_).max() < 1e-8
    for a, b in zip(full.loadings_, red.loadings_):
        assert np.abs(np.abs(a) - np.abs(b)).max() < 1e-6
    low = mcca_restated.MCCARestated(n_components=4, regs=0.5, signal_ranks=[3, 3, 3]).fit(views)
    assert all(np.linalg.matrix_rank(l) <= 3 for l in low.loadings_)
    with pytest.raises(ValueError):
        mcca_restated.MCCARestated(n_components=10, regs=0.5, signal_ranks=[3, 3, 3]).fit(views)


def test_mcca_view_permutation_invariance():
    views = _views(3)
    a = mcca_restated.MCCARestated(n_components=4, regs=0.5).fit(views)
    b = mcca_restated.MCCARestated(n_components=4, regs=0.5).fit(views[::-1])
    assert np.abs(a.evals_ - b.evals_).max() < 1e-9
    assert np.abs(np.abs(a.loadings_[0]) - np.abs(b.loadings_[2])).max() < 1e-6


def test_exact_svm_matches_liblinear_primal():
    from sklearn.svm import LinearSVC
    rng = np.random.default_rng(0)
    X = rng.standard_normal((200, 12)) * np.linspace(40, 2, 12)
    y = rng.integers(0, 4, 200)
    cls, W = svm_exact.solve_ovr(X, y)
    ref = LinearSVC(dual=False, C=1.0, tol=1e-12, max_iter=100000).fit(X, y)
    Wref = np.hstack([ref.coef_, ref.intercept_[:, None]])
    assert np.abs(W - Wref).max() <= 1e-6 * np.abs(Wref).max()
    X1 = np.hstack([X, np.ones((200, 1))])
    for c, w in zip(cls, W):
        g = svm_exact.gradient(w, X1, np.where(y == c, 1.0, -1.0), 1.0)
        assert np.abs(g).max() < 1e-6


def test_liblinear_dual_cd_does_not_converge_on_unscaled_scores():
    """Documents why the oracle decoder is the primal liblinear solver: on PCA-score-like
    features the dual CD of LinearSVC(dual=True) stops at max_iter far from the optimum."""
    from sklearn.exceptions import ConvergenceWarning
    from sklearn.svm import LinearSVC
    rng = np.random.default_rng(1)
    X = rng.standard_normal((250, 20)) * np.linspace(70, 25, 20)
    y = (X @ rng.standard_normal(20) + 60 * rng.standard_normal(250) > 0).astype(int)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        dual = LinearSVC(dual=True, C=1.0, max_iter=1000, random_state=0).fit(X, y)
    assert any(issubclass(x.category, ConvergenceWarning) for x in w)
    prim = LinearSVC(dual=False, C=1.0, tol=1e-10, max_iter=100000).fit(X, y)
    rel = np.linalg.norm(dual.coef_ - prim.coef_) / np.linalg.norm(prim.coef_)
    assert rel > 0.05


@pytest.mark.parametrize('name', ['cca_p3_ragged', 'none_p3_ragged', 'mcca_p3_ragged',
                                  'cca_p3_svc_rbf', 'cca_p3_d60', 'jointpca_p3_d60', 'cca_real_shapes_t7'])
def test_cpu_port_reproduces_reference_golden(name):
    """oracle/pipeline_port.py (used on the GPU box, where /root/reference is absent) against
    the outputs the UNMODIFIED reference produced here (tests/golden/make_golden.py)."""
    import make_golden
    cfg = make_golden.CONFIGS[name]
    pts, folds = make_golden.build_inputs(cfg)
    g = np.load(os.path.join(HERE, 'golden', name + '.npz'))
    for f in range(2 if len(pts) <= 3 and pts[0][0].shape[2] < 100 else 1):
        tr, te = folds[f]
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            yp, k2 = pipeline_port.run_fold(pts[0], pts[1:], tr, te, method=cfg['method'],
                                            n_comp=cfg.get('n_comp'), regs=cfg.get('regs', 0.5),
                                            pca_var=cfg.get('pca_var', 0.8),
                                            decoder=cfg.get('svm', 'linear').replace('primal', 'linear'),
                                            class_weight='balanced' if cfg.get('svm') == 'svc_rbf' else None)
        assert k2 == int(g['k2'][f])
        assert np.array_equal(yp, g['y_pred_%d' % f])


@pytest.mark.skipif(not reference_path.available(), reason='reference tree not present')
def test_reference_classes_match_port_live():
    """With /root/reference importable: its AlignCCA / cnd_avg against the port, same inputs."""
    from cross_patient_speech_decoding_b200 import synthetic
    ref = reference_path.load()
    (Xa, _, ya), (Xb, _, yb) = [synthetic.make_patient(p, n_trials=50, n_time=30, n_chan=10)
                                for p in range(2)]
    al = ref.AlignCCA()
    al.fit(Xa, Xb, ya, yb)
    Ma, Mb, rho = pipeline_port.cca_fit(Xa, Xb, ya, yb)
    assert np.abs(al.canon_corrs - rho).max() < 1e-12
    s = pipeline_port.labels_as_str(ya)
    assert np.abs(ref.utils.cnd_avg(Xa, s) - pipeline_port.condition_average(Xa, s)).max() < 1e-12
    am = ref.AlignMCCA(n_components=6, regs=0.5, pca_var=0.8)
    am.fit([Xa, Xb], [ya, yb])
    pm = pipeline_port.mcca_fit([Xa, Xb], [ya, yb], 6, 0.5, 0.8)
    assert np.abs(am.mcca.evals_ - pm.evals_).max() < 1e-10


def test_joint_pca_port_matches_reference_golden():
    """oracle/pipeline_port.joint_pca_* against the output of the reference's own JointPCA
    (tests/golden/jointpca_p3_ragged.npz, made by make_golden_jointpca.py)."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, 'golden'))
    import make_golden
    from oracle import pipeline_port as port
    g = np.load(os.path.join(here, 'golden', 'jointpca_p3_ragged.npz'))
    pts, _ = make_golden.build_inputs(make_golden.CONFIGS['mcca_p3_ragged'])
    W = port.joint_pca_fit([p[0] for p in pts], [p[2] for p in pts], int(g['n_comp']))
    for v, p in enumerate(pts):
        Wref = g['W_full_%d' % v]
        assert np.abs(W[v] - Wref).max() <= 1e-8 * np.abs(Wref).max()
        Z = port.joint_pca_transform(W[v], p[0][:4])
        assert np.abs(Z - g['Z_full_%d' % v]).max() <= 1e-8 * np.abs(g['Z_full_%d' % v]).max()


@pytest.mark.parametrize('kernel,balanced,ncls', [('rbf', True, 4), ('rbf', False, 2), ('linear', False, 3)])
def test_svc_smo_restatement_matches_libsvm(kernel, balanced, ncls):
    """oracle/svc_smo.py (the algorithm csrc/svc.cu follows) against sklearn.svm.SVC = libsvm,
    the class the reference scripts instantiate (scripts/aligned_decode_svm_ncv.py:313-317)."""
    from sklearn.svm import SVC
    from oracle import svc_smo
    rng = np.random.default_rng(5)
    y = rng.integers(0, ncls, 150) + 1
    cent = rng.standard_normal((ncls + 1, 6)) * 0.9
    X = cent[y] + rng.standard_normal((150, 6))
    Z = cent[rng.integers(1, ncls + 1, 60)] + rng.standard_normal((60, 6))
    ref = SVC(kernel=kernel, class_weight='balanced' if balanced else None, tol=1e-6,
              decision_function_shape='ovo').fit(X, y)
    m = svc_smo.fit_ovo(X, y, kernel=kernel, balanced=balanced, tol=1e-6)
    if kernel == 'rbf':
        assert abs(m['gamma'] - ref._gamma) <= 1e-12 * ref._gamma
    d_ref = ref.decision_function(Z)
    d = svc_smo.decision_ovo(m, Z)
    d = -d[:, 0] if ncls == 2 else d                  # sklearn flips the binary sign
    assert np.abs(d - d_ref).max() <= 2e-4 * max(1.0, np.abs(d_ref).max())
    assert np.array_equal(svc_smo.predict_ovo(m, Z), ref.predict(Z))
    rho = np.array([p['rho'] for p in m['pairs']])
    assert np.abs((-rho if ncls > 2 else rho) - ref.intercept_).max() < 1e-4


@pytest.mark.skipif(not reference_path.available(), reason='reference tree not present')
def test_port_jointpca_and_svc_match_reference_classes_live():
    """With /root/reference importable: crossPtDecoder_jointDimRed + JointPCA with a variance
    fraction (what the script's set_params produces) and the scripts' own SVC decoder, against
    the port's run_fold on the same fold."""
    from sklearn.decomposition import PCA
    from sklearn.pipeline import make_pipeline
    from sklearn.svm import SVC
    from cross_patient_speech_decoding_b200 import synthetic
    ref = reference_path.load()
    pts = [synthetic.make_patient(p, n_trials=60, n_time=20, n_chan=12, noise=0.8) for p in range(3)]
    Xt, yt, yat = pts[0]
    tr, te = np.arange(0, 45), np.arange(45, 60)
    clf = make_pipeline(ref.DimRedReshape(PCA, n_components=0.8), SVC(kernel='rbf', class_weight='balanced'))
    m = ref.decoders.crossPtDecoder_jointDimRed(pts[1:], clf, ref.JointPCA, n_comp=0.9)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m.fit(Xt[tr], yt[tr], y_align=yat[tr])
        want = m.predict(Xt[te])
        got, k2 = pipeline_port.run_fold(pts[0], pts[1:], tr, te, method='jointpca', n_comp=0.9,
                                         decoder='svc_rbf', class_weight='balanced')
    assert k2 == clf.steps[0][1].transformer.n_components_
    assert np.array_equal(got, want)


@pytest.mark.skipif(not reference_path.available(), reason='reference tree not present')
def test_port_trial_subselect_matches_reference_live():
    """oracle/pipeline_port.cca_fit_trial against the reference's AlignCCA(type='trial') under the
    same numpy seed: identical trial draws, canonical correlations and maps."""
    from cross_patient_speech_decoding_b200 import synthetic
    ref = reference_path.load()
    (Xa, _, ya), (Xb, _, yb) = [synthetic.make_patient(p, n_trials=70, n_time=20, n_chan=10) for p in range(2)]
    np.random.seed(17)
    al = ref.AlignCCA(type='trial')
    al.fit(Xa, Xb, ya, yb)
    np.random.seed(17)
    Ma, Mb, rho = pipeline_port.cca_fit_trial(Xa, Xb, ya, yb)
    assert np.abs(rho - al.canon_corrs).max() < 1e-12
    assert np.abs(np.abs(Ma) - np.abs(al.M_a)).max() <= 1e-9 * np.abs(al.M_a).max()
    assert np.abs(np.abs(Mb) - np.abs(al.M_b)).max() <= 1e-9 * np.abs(al.M_b).max()
