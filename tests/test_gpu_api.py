"""The sklearn-style drop-in classes (same names / signatures / errors as the reference)
against float64 numpy statements of the reference arithmetic and against golden outputs."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))


@pytest.fixture(scope='module')
def pkg(lib_built):
    import cross_patient_speech_decoding_b200 as p
    return p


def _patients(n, **kw):
    from cross_patient_speech_decoding_b200 import synthetic
    base = dict(n_trials=60, n_time=40, n_chan=32)
    base.update(kw)
    return [synthetic.make_patient(p, **base) for p in range(n)]


def test_cnd_avg_and_group_conditions(pkg):
    from cross_patient_speech_decoding_b200.alignment import alignment_utils as au
    from oracle import pipeline_port as port
    pts = _patients(3)
    X, _, ya = pts[0]
    got = au.cnd_avg(X, au.label2str(ya))
    ref = port.condition_average(X, port.labels_as_str(ya))
    assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-5
    g = au.extract_group_conditions([p[0] for p in pts], [p[2] for p in pts])
    r = port.shared_condition_averages([p[0] for p in pts], [p[2] for p in pts])
    for a, b in zip(g, r):
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-5
    assert list(au.label_seq2str(np.array([[1, 2, 3], [9, 9, 1]]))) == ['123', '991']
    assert list(au.phon_to_artic_seq(np.array([1, 5, 9]))) == [1, 3, 4]


def test_pca_matches_sklearn_tall_and_wide(pkg):
    from sklearn.decomposition import PCA as SkPCA
    from cross_patient_speech_decoding_b200.decomposition.PCA import PCA
    rng = np.random.default_rng(0)
    tall = rng.standard_normal((3000, 40)) @ np.diag(np.linspace(3, 0.2, 40)) + 1.5
    wide = rng.standard_normal((150, 900)) * np.linspace(4, 0.5, 900) - 0.5
    # > 256 on the eigen side: the top-k subspace solver path of the PCA class
    wide2 = rng.standard_normal((400, 900)) * (1.0 / (1.0 + np.arange(900) / 15.0)) + 0.3
    tall2 = rng.standard_normal((1500, 300)) * (1.0 / (1.0 + np.arange(300) / 10.0)) - 1.0
    for X, nc in [(tall, 0.9), (tall, 7), (wide, 0.8), (wide, 12), (tall, None), (wide2, 0.8),
                  (wide2, 12), (tall2, 10), (tall2, 0.9)]:
        # svd_solver='full': sklearn's 'auto' picks the randomized solver for the wide int case
        ours, ref = PCA(n_components=nc).fit(X), SkPCA(n_components=nc, svd_solver='full').fit(X)
        assert ours.n_components_ == ref.n_components_
        k = ref.n_components_
        assert np.abs(ours.explained_variance_ratio_ - ref.explained_variance_ratio_).max() < 1e-5
        assert np.abs(ours.mean_ - ref.mean_).max() < 1e-5
        kk = min(k, 5)           # leading components are well separated: compare with signs
        assert np.abs(ours.components_[:kk] - ref.components_[:kk]).max() < 2e-3
        Z, Zr = ours.transform(X[:20]), ref.transform(X[:20])
        assert np.abs(Z[:, :kk] - Zr[:, :kk]).max() <= 2e-3 * np.abs(Zr).max()


def test_nocenter_pca(pkg):
    from cross_patient_speech_decoding_b200.decomposition.NoCenterPCA import NoCenterPCA
    rng = np.random.default_rng(1)
    X = rng.standard_normal((500, 30)) * np.linspace(5, 0.5, 30) + 2.0
    _, S, Vt = np.linalg.svd(X, full_matrices=False)
    cum = np.cumsum(S ** 2) / np.sum(S ** 2)
    p = NoCenterPCA(n_components=0.9).fit(X)
    k = int(np.argmax(cum >= 0.9) + 1)
    assert p.components_.shape == (30, k)
    assert np.abs(p.explained_variance_ - S ** 2).max() <= 1e-5 * S[0] ** 2
    assert np.abs(np.abs(p.transform(X)) - np.abs(X @ Vt[:k].T)).max() <= 1e-3 * np.abs(X).max()
    with pytest.raises(ValueError, match='PCA must be fit before transforming data.'):
        NoCenterPCA(3).transform(X)
    assert NoCenterPCA(n_components=None).fit(X).components_.shape == (30, 30)


def test_align_cca_class(pkg):
    from cross_patient_speech_decoding_b200.alignment.AlignCCA import AlignCCA
    from oracle import pipeline_port as port
    pts = _patients(2, n_chan=12)
    (Xa, _, ya), (Xb, _, yb) = pts
    al = AlignCCA()
    with pytest.raises(RuntimeError, match=r'Must call fit\(\) before transforming data.'):
        al.transform(Xb)
    al.fit(Xa, Xb, ya, yb)
    Ma, Mb, rho = port.cca_fit(Xa, Xb, ya, yb)
    assert al.canon_corrs.shape == rho.shape
    assert np.abs(al.canon_corrs - rho).max() < 1e-4
    ref = Xb @ Mb @ np.linalg.pinv(Ma)
    got = al.transform(Xb)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-3 * np.abs(ref).max()
    sh = AlignCCA(return_space='shared')
    sh.fit(Xa, Xb, ya, yb)
    za, zb = sh.transform([Xa, Xb])
    assert za.shape[:2] == Xa.shape[:2] and zb.shape[-1] == za.shape[-1]
    ab = AlignCCA(return_space='a_to_b')
    ab.fit(Xa, Xb, ya, yb)
    ref_ab = Xa @ Ma @ np.linalg.pinv(Mb)
    assert np.abs(ab.transform(Xa) - ref_ab).max() <= 2e-3 * np.abs(ref_ab).max()
    with pytest.raises(ValueError, match='type must be "class" or "trial".'):
        AlignCCA(type='bogus').fit(Xa, Xb, ya, yb)


def test_align_cca_trial_type_seeded(pkg):
    """AlignCCA(type='trial') (AlignCCA.py:186-232): under the same numpy seed the class draws the
    reference's trial subsets (np.random.permutation per shared class, A then B) -- canonical
    correlations within 1e-4 and the b->a transform against the CPU port of the reference (pinned
    live against the reference class in tests/test_oracle.py)."""
    from cross_patient_speech_decoding_b200.alignment.AlignCCA import AlignCCA
    from oracle import pipeline_port as port
    (Xa, _, ya), (Xb, _, yb) = _patients(2, n_trials=90, n_chan=14)
    for seed in (3, 8):
        np.random.seed(seed)
        al = AlignCCA(type='trial')
        al.fit(Xa, Xb, ya, yb)
        state = np.random.get_state()[1].copy()
        np.random.seed(seed)
        Ma, Mb, rho = port.cca_fit_trial(Xa, Xb, ya, yb)
        assert np.array_equal(state, np.random.get_state()[1])      # same number of draws
        assert al.canon_corrs.shape == rho.shape
        assert np.abs(al.canon_corrs - rho).max() < 1e-4
        ref = Xb @ Mb @ np.linalg.pinv(Ma)
        got = al.transform(Xb)
        assert np.abs(got - ref).max() <= 2e-3 * np.abs(ref).max()


def test_cca_rank_deficient_latents_raise(pkg):
    """Rank-deficient class-averaged latents (a duplicated dimension): the reference truncates to
    matrix_rank (AlignCCA.py:263-265); the Gram-form solver reports it instead of returning
    garbage."""
    from cross_patient_speech_decoding_b200 import ops
    from cross_patient_speech_decoding_b200.alignment.AlignCCA import CCA_align
    rng = np.random.default_rng(0)
    La = rng.standard_normal((6, 400))
    Lb = rng.standard_normal((5, 400))
    La[5] = La[2]                                   # exact duplicate -> singular scatter
    with pytest.raises(np.linalg.LinAlgError):
        CCA_align(La, Lb)
    Lc = La - La.mean(axis=1, keepdims=True)
    Ld = Lb - Lb.mean(axis=1, keepdims=True)
    out = ops.cca_solve(Lc @ Lc.T, Ld @ Ld.T, Lc @ Ld.T, check_rank=False)
    assert out['info'][1] == 1


def test_align_mcca_class(pkg):
    from cross_patient_speech_decoding_b200.alignment.AlignMCCA import AlignMCCA, n_components_var
    from oracle import pipeline_port as port
    pts = _patients(3)
    Xs, ys = [p[0] for p in pts], [p[2] for p in pts]
    al = AlignMCCA(n_components=8, regs=0.5, pca_var=0.8)
    with pytest.raises(RuntimeError):
        al.transform(Xs)
    out = al.fit_transform(Xs, ys)
    model = port.mcca_fit(Xs, ys, 8, 0.5, 0.8)
    assert len(al.mcca.loadings_) == 3
    assert np.abs(al.mcca.evals_ - model.evals_).max() <= 2e-4 * np.abs(model.evals_).max()
    for v in range(3):
        ref = port.mcca_transform(model, Xs[v], v)
        assert out[v].shape == ref.shape
        # same deterministic sign convention -> directly comparable
        assert np.abs(out[v] - ref).max() <= 2e-2 * np.abs(ref).max()
    one = al.transform(Xs[1], idx=1)
    assert np.abs(one - out[1]).max() < 1e-6
    with pytest.raises(IndexError, match='Input idx is greater than the number of learned'):
        al.transform(Xs[0], idx=3)
    X2 = Xs[0].reshape(-1, Xs[0].shape[-1])
    assert n_components_var(X2, 0.8) == port.signal_rank(X2, 0.8)


def test_cross_pt_decoders_drop_in(pkg):
    """Reference-style usage: make_pipeline(DimRedReshape(PCA), LinearSVC) injected into the
    crossPtDecoder classes; predictions equal the CPU port's on a small fold."""
    from sklearn.pipeline import make_pipeline
    from cross_patient_speech_decoding_b200.alignment.AlignCCA import AlignCCA
    from cross_patient_speech_decoding_b200.alignment.AlignMCCA import AlignMCCA
    from cross_patient_speech_decoding_b200.decoders.cross_pt_decoders import (
        crossPtDecoder_mcca, crossPtDecoder_sepAlign, crossPtDecoder_sepDimRed)
    from cross_patient_speech_decoding_b200.decomposition.DimRedReshape import DimRedReshape
    from cross_patient_speech_decoding_b200.decomposition.PCA import PCA
    from cross_patient_speech_decoding_b200.svm import LinearSVC
    from oracle import pipeline_port as port
    from cross_patient_speech_decoding_b200.folds import cv_splits
    pts = _patients(3, n_trials=120)
    Xt, yt, yat = pts[0]
    np.random.seed(12)
    folds = cv_splits(yt, 4)                      # 120 held-out labels per decoder class
    tr, te = folds[0]
    for cls, kw, method, nc in [
            (crossPtDecoder_sepAlign, dict(aligner=AlignCCA, n_comp=0.9), 'cca', 0.9),
            (crossPtDecoder_sepDimRed, dict(n_comp=0.9), 'none', 0.9),
            (crossPtDecoder_mcca, dict(aligner=AlignMCCA, n_comp=8, regs=0.5, pca_var=0.8),
             'mcca', 8)]:
        clf = make_pipeline(DimRedReshape(PCA, n_components=0.8), LinearSVC())
        m = cls(pts[1:], clf, **kw)
        fitted = m.fit(Xt[tr], yt[tr], y_align=yat[tr]) if method != 'none' else \
            m.fit(Xt[tr], yt[tr])
        assert fitted is clf                      # reference quirk: fit returns the decoder
        yp = m.predict(Xt[te])
        ref, k2 = port.run_fold(pts[0], pts[1:], tr, te, method=method, n_comp=nc)
        assert clf.steps[0][1].transformer.n_components_ == k2
        same, tot = int((yp == ref).sum()), len(ref)
        assert 0.0 <= m.score(Xt[te], yt[te]) <= 1.0
        for tr2, te2 in folds[1:]:                # the remaining folds: fresh estimators, as the
            clf2 = make_pipeline(DimRedReshape(PCA, n_components=0.8), LinearSVC())   # scripts do
            m2 = cls(pts[1:], clf2, **kw)
            m2.fit(Xt[tr2], yt[tr2], y_align=yat[tr2]) if method != 'none' else m2.fit(Xt[tr2], yt[tr2])
            ref2, k22 = port.run_fold(pts[0], pts[1:], tr2, te2, method=method, n_comp=nc)
            assert clf2.steps[0][1].transformer.n_components_ == k22
            same += int((m2.predict(Xt[te2]) == ref2).sum())
            tot += len(ref2)
        assert tot >= 100 and same / tot >= 0.99, (method, same, tot)
    m = crossPtDecoder_mcca(pts[1:], make_pipeline(DimRedReshape(PCA, 0.8), LinearSVC()),
                            AlignMCCA, n_comp=8, regs=0.5, pca_var=0.8)
    m.fit(Xt[tr], yt[tr], y_align=yat[tr])
    with pytest.raises(TypeError):                # second fit: aligner is now an instance
        m.fit(Xt[tr], yt[tr], y_align=yat[tr])
    # sklearn plumbing used by the reference scripts
    m2 = crossPtDecoder_sepAlign(pts[1:], make_pipeline(DimRedReshape(PCA), LinearSVC()), AlignCCA)
    m2.set_params(**{'n_comp': 0.9, 'decoder__dimredreshape__n_components': 0.8})
    assert m2.get_params()['decoder__dimredreshape__n_components'] == 0.8


def test_cv_align_decode_public_api(pkg):
    import make_golden
    cfg = make_golden.CONFIGS['cca_p3_ragged']
    pts, folds = make_golden.build_inputs(cfg)
    g = np.load(os.path.join(HERE, 'golden', 'cca_p3_ragged.npz'))
    out = pkg.cv_align_decode(pts[0], pts[1:], folds, method='cca', n_comp=0.9)
    yp = np.concatenate(out['y_pred'])
    yr = np.concatenate([g['y_pred_%d' % f] for f in range(len(folds))])
    assert np.mean(yp == yr) >= 0.99
    assert out['h2d_bytes'] > 0 and out['d2h_bytes'] > 0


def test_joint_pca_class_with_more_than_128_electrodes(pkg):
    """JointPCA read-in matrices of a patient with 150 electrodes (the real recordings reach 201):
    the normal equations factor in the workspace-backed Cholesky; against the float64 port of
    JointPCA.get_joint_PCA_transforms (pinv(X_p) @ latent)."""
    from cross_patient_speech_decoding_b200 import ops, synthetic
    from cross_patient_speech_decoding_b200.alignment.JointPCA import JointPCA
    from oracle import pipeline_port as port
    pts = [synthetic.make_patient(p, n_trials=n, n_time=30, n_chan=c) for p, n, c in ((0, 80, 150), (1, 90, 40))]
    Xs, yal = [p[0] for p in pts], [p[2] for p in pts]
    jp = JointPCA(n_components=8)
    Z = jp.fit_transform(Xs, yal)
    Wref = port.joint_pca_fit(Xs, yal, 8)
    for v in range(2):
        W = jp.transforms[v]
        sgn = np.sign(np.sum(W * Wref[v], axis=0))
        assert np.abs(W * sgn - Wref[v]).max() <= 5e-3 * np.abs(Wref[v]).max()
        assert Z[v].shape == Xs[v].shape[:-1] + (8,)
    # the solver alone, 201 columns
    rng = np.random.default_rng(0)
    X = rng.standard_normal((900, 201))
    Y = rng.standard_normal((900, 7))
    W, st = ops.lstsq_gram(X, Y)
    ref = np.linalg.lstsq(X.astype(np.float32).astype(np.float64), Y.astype(np.float32).astype(np.float64),
                          rcond=None)[0]
    assert st == 0 and np.abs(W - ref).max() <= 1e-5 * np.abs(ref).max()


def test_joint_pca_matches_reference_golden(pkg):
    """alignment.JointPCA.JointPCA (GPU) against the reference class's stored output: read-in
    matrices, transformed trials, API errors; and crossPtDecoder_jointDimRed end to end."""
    import make_golden
    from cross_patient_speech_decoding_b200.alignment.JointPCA import JointPCA
    from cross_patient_speech_decoding_b200.decoders.cross_pt_decoders import crossPtDecoder_jointDimRed
    from cross_patient_speech_decoding_b200.decomposition.DimRedReshape import DimRedReshape
    from cross_patient_speech_decoding_b200.decomposition.PCA import PCA
    from cross_patient_speech_decoding_b200.svm import LinearSVC
    from sklearn.pipeline import make_pipeline
    g = np.load(os.path.join(HERE, 'golden', 'jointpca_p3_ragged.npz'))
    pts, folds = make_golden.build_inputs(make_golden.CONFIGS['mcca_p3_ragged'])
    Xs, yal = [p[0] for p in pts], [p[2] for p in pts]
    nc = int(g['n_comp'])
    jp = JointPCA(n_components=nc)
    with pytest.raises(RuntimeError):
        jp.transform(Xs)
    Z = jp.fit_transform(Xs, yal)
    assert isinstance(Z, tuple) and len(Z) == 3 and len(jp.transforms) == 3
    for v in range(3):
        Wref, Zref = g['W_full_%d' % v], g['Z_full_%d' % v]
        assert jp.transforms[v].shape == Wref.shape
        # the transformed data is what downstream stages consume: compare it tightly, and the
        # read-in matrices through the subspace they span
        assert np.abs(Z[v][:4] - Zref).max() <= 2e-4 * np.abs(Zref).max()
        assert np.abs(jp.transforms[v] - Wref).max() <= 5e-3 * np.abs(Wref).max()
    zs = jp.transform(Xs[1][:3], idx=1)
    assert np.abs(zs - g['Zsingle_full']).max() <= 2e-4 * np.abs(g['Zsingle_full']).max()
    with pytest.raises(IndexError):
        jp.transform(Xs[0], idx=3)
    # decoder
    Xt, yt, yat = pts[0]
    agree = tot = 0
    for f, (tr, te) in enumerate(folds):
        clf = make_pipeline(DimRedReshape(PCA, n_components=0.8), LinearSVC())
        model = crossPtDecoder_jointDimRed(pts[1:], clf, JointPCA, n_comp=nc)
        model.fit(Xt[tr], yt[tr], y_align=yat[tr])
        yp = model.predict(Xt[te])
        assert clf.named_steps['dimredreshape'].transformer.n_components_ == int(g['k2_%d' % f])
        agree += int((yp == g['y_pred_%d' % f]).sum())
        tot += len(te)
    assert agree / tot >= 0.99


def test_pt_corr_matches_scipy(pkg):
    """alignment.metrics.pt_corr / pt_corr_multi against scipy.stats.pearsonr (what the reference
    calls per condition, alignment/metrics.py:41-68)."""
    from scipy.stats import pearsonr
    from cross_patient_speech_decoding_b200.alignment.metrics import pt_corr, pt_corr_multi
    rng = np.random.default_rng(2)
    a = rng.standard_normal((7, 30, 6))
    b = 0.6 * a + 0.8 * rng.standard_normal((7, 30, 6))
    b[3] = -a[3]
    r, p = pt_corr(a, b, p_vals=True)
    ref = [pearsonr(a[c].ravel(), b[c].ravel()) for c in range(7)]
    assert np.abs(r - np.array([x[0] for x in ref])).max() < 1e-12
    assert np.allclose(p, np.array([x[1] for x in ref]), rtol=1e-9, atol=1e-300)
    rs = pt_corr_multi(a, [b, a])
    assert np.allclose(rs[1], 1.0) and np.abs(rs[0] - r).max() == 0
    with pytest.raises(ValueError):
        pt_corr(a, b[:, :10])


def test_fused_predictor_matches_class_predict(pkg):
    """decoders.fused_predict.FusedPredictor (one kernel per call) against the three-stage
    crossPtDecoder.predict of the same fitted model, for every decoder class."""
    from sklearn.pipeline import make_pipeline
    from cross_patient_speech_decoding_b200.alignment.AlignCCA import AlignCCA
    from cross_patient_speech_decoding_b200.alignment.AlignMCCA import AlignMCCA
    from cross_patient_speech_decoding_b200.alignment.JointPCA import JointPCA
    from cross_patient_speech_decoding_b200.decoders.cross_pt_decoders import (
        crossPtDecoder_jointDimRed, crossPtDecoder_mcca, crossPtDecoder_sepAlign,
        crossPtDecoder_sepDimRed)
    from cross_patient_speech_decoding_b200.decoders.fused_predict import FusedPredictor
    from cross_patient_speech_decoding_b200.decomposition.DimRedReshape import DimRedReshape
    from cross_patient_speech_decoding_b200.decomposition.PCA import PCA
    from cross_patient_speech_decoding_b200.svm import LinearSVC
    pts = _patients(3, n_trials=80)
    Xt, yt, yat = pts[0]
    tr, te = np.arange(0, 60), np.arange(60, 80)
    for cls, kw, fitkw in [
            (crossPtDecoder_mcca, dict(aligner=AlignMCCA, n_comp=8, regs=0.5, pca_var=0.8), True),
            (crossPtDecoder_sepAlign, dict(aligner=AlignCCA, n_comp=0.9), True),
            (crossPtDecoder_sepDimRed, dict(n_comp=0.9), False),
            (crossPtDecoder_jointDimRed, dict(joint_dr_method=JointPCA, n_comp=10), True)]:
        m = cls(pts[1:], make_pipeline(DimRedReshape(PCA, n_components=0.8), LinearSVC()), **kw)
        m.fit(Xt[tr], yt[tr], y_align=yat[tr]) if fitkw else m.fit(Xt[tr], yt[tr])
        fp = FusedPredictor(m)
        ref = m.predict(Xt[te])
        got, dec = fp.predict(Xt[te], return_decision=True)
        assert got.dtype == ref.dtype and np.mean(got == ref) >= 0.95, (cls.__name__, got, ref)
        Xp = m.preprocess_test(Xt[te])
        dref = m.decoder.decision_function(Xp)
        assert np.abs(dec - dref).max() <= 2e-3 * max(1.0, np.abs(dref).max()), cls.__name__
        assert np.array_equal(fp.predict(Xt[te][:1]), got[:1])          # batch of one
    with pytest.raises(ValueError):
        fp.predict(Xt[te][:, :5])
    # the scripts' own decoder behind the fused front end (scores kernel + C-SVC vote kernel)
    from cross_patient_speech_decoding_b200.svm import SVC
    m = crossPtDecoder_mcca(pts[1:], make_pipeline(DimRedReshape(PCA, n_components=0.8),
                                                   SVC(kernel='rbf', class_weight='balanced')),
                            AlignMCCA, n_comp=8, regs=0.5, pca_var=0.8)
    m.fit(Xt[tr], yt[tr], y_align=yat[tr])
    fp = FusedPredictor(m)
    assert np.mean(fp.predict(Xt[te]) == m.predict(Xt[te])) >= 0.95
    assert np.array_equal(fp.predict(Xt[te][3:4]), fp.predict(Xt[te])[3:4])


def test_realtime_hooks(pkg):
    """realtime_sim.reduce_to_latent_space / align_to_target against sklearn PCA and the CPU port
    of CCA_align (what the reference's hooks call, realtime_datamodule.py:813-894)."""
    import torch
    from sklearn.decomposition import PCA as SkPCA
    from cross_patient_speech_decoding_b200.alignment.AlignCCA import AlignCCA
    from cross_patient_speech_decoding_b200.realtime_sim import align_to_target, reduce_to_latent_space
    pts = _patients(2, n_trials=60)
    Xa, Xb = torch.tensor(pts[0][0]), torch.tensor(pts[1][0])
    ra, pca_a = reduce_to_latent_space(Xa, n_components=6)
    ref = SkPCA(n_components=6).fit(pts[0][0].reshape(-1, pts[0][0].shape[-1]))
    assert tuple(ra.shape) == (60, Xa.shape[1], 6) and ra.dtype == torch.float32
    want = ref.transform(pts[0][0].reshape(-1, pts[0][0].shape[-1])).reshape(60, -1, 6)
    assert np.abs(ra.numpy() - want).max() <= 2e-3 * np.abs(want).max()
    rb2, _ = reduce_to_latent_space(Xa[:5], pca=pca_a)                 # transform-only branch
    assert np.abs(rb2.numpy() - ra.numpy()[:5]).max() <= 1e-4 * np.abs(want).max()
    # variance threshold that keeps <= 5 components -> the 30-component re-fit
    r30, p30 = reduce_to_latent_space(Xa, n_components=0.2)
    assert p30.n_components_ == 30 and r30.shape[-1] == 30
    rb, _ = reduce_to_latent_space(Xb, n_components=6)
    ya, yb = torch.tensor(pts[0][2]), torch.tensor(pts[1][2])
    out = align_to_target(AlignCCA, ra, rb, ya, yb)
    chk = AlignCCA()
    chk.fit(ra.numpy(), rb.numpy(), pts[0][2], pts[1][2])
    assert tuple(out.shape) == (60, Xb.shape[1], 6)
    assert np.abs(out.numpy() - chk.transform(rb.numpy())).max() <= 1e-5 * np.abs(out.numpy()).max() + 1e-6


def test_nested_search_with_sklearn_searchcv(pkg):
    """The nested-CV pattern of scripts/aligned_decode_svm_ncv.py:388-405 (search with
    refit=False over the decoder's hyper-parameters, extra fit kwarg y_align, then set_params +
    fit) with sklearn's own GridSearchCV driving the GPU estimators (BayesSearchCV needs the
    absent scikit-optimize; the script lists GridSearchCV / RandomizedSearchCV as the
    alternatives)."""
    from sklearn.model_selection import GridSearchCV, StratifiedKFold
    from sklearn.pipeline import make_pipeline
    from cross_patient_speech_decoding_b200.alignment.AlignCCA import AlignCCA
    from cross_patient_speech_decoding_b200.decoders.cross_pt_decoders import crossPtDecoder_sepAlign
    from cross_patient_speech_decoding_b200.decomposition.DimRedReshape import DimRedReshape
    from cross_patient_speech_decoding_b200.decomposition.PCA import PCA
    from cross_patient_speech_decoding_b200.svm import SVC
    pts = _patients(2, n_trials=60)
    Xt, yt, yat = pts[0]
    clf = make_pipeline(DimRedReshape(PCA), SVC(kernel='rbf', class_weight='balanced'))
    model = crossPtDecoder_sepAlign(pts[1:], clf, AlignCCA)
    grid = {'n_comp': [0.8, 0.9], 'decoder__dimredreshape__n_components': [0.6, 0.8]}
    search = GridSearchCV(model, grid, cv=StratifiedKFold(2, shuffle=True, random_state=0), refit=False)
    search.fit(Xt, yt, y_align=yat)
    assert set(search.best_params_) == set(grid)
    assert 0.0 <= search.best_score_ <= 1.0 and len(search.cv_results_['params']) == 4
    model.set_params(**search.best_params_)
    model.fit(Xt, yt, y_align=yat)
    assert model.predict(Xt[:5]).shape == (5,)


def test_batched_search_matches_gridsearchcv(pkg):
    """search_align_decode (all candidates x inner folds streamed through the engine, patients
    uploaded once) against sklearn's GridSearchCV around the drop-in estimators on the same
    inner folds: same mean accuracies per candidate, same winner."""
    from sklearn.model_selection import GridSearchCV, StratifiedKFold
    from sklearn.pipeline import make_pipeline
    from cross_patient_speech_decoding_b200.alignment.AlignCCA import AlignCCA
    from cross_patient_speech_decoding_b200.decoders.cross_pt_decoders import crossPtDecoder_sepAlign
    from cross_patient_speech_decoding_b200.decomposition.DimRedReshape import DimRedReshape
    from cross_patient_speech_decoding_b200.decomposition.PCA import PCA
    from cross_patient_speech_decoding_b200.svm import LinearSVC
    pts = _patients(3, n_trials=72)
    Xt, yt, yat = pts[0]
    cv = StratifiedKFold(3, shuffle=True, random_state=1)
    folds = list(cv.split(np.zeros((len(yt), 1)), yt))
    cands = [dict(n_comp=a, decoder_var=b) for a in (0.8, 0.9) for b in (0.6, 0.8)]
    out = pkg.search_align_decode(pts[0], pts[1:], cands, folds, method='cca', depth=2)
    model = crossPtDecoder_sepAlign(pts[1:], make_pipeline(DimRedReshape(PCA), LinearSVC()), AlignCCA)
    grid = {'n_comp': [0.8, 0.9], 'decoder__dimredreshape__n_components': [0.6, 0.8]}
    gs = GridSearchCV(model, grid, cv=folds, refit=False).fit(Xt, yt, y_align=yat)
    ref = {(p['n_comp'], p['decoder__dimredreshape__n_components']): s
           for p, s in zip(gs.cv_results_['params'], gs.cv_results_['mean_test_score'])}
    got = {(c['n_comp'], c['decoder_var']): s for c, s in zip(cands, out['scores'])}
    for k in ref:
        assert abs(ref[k] - got[k]) <= 0.03, (k, ref[k], got[k])
    assert out['best_params'] == cands[out['best_index']] and len(out['y_pred']) == 4


def test_fused_predictor_matches_oracle_port(pkg):
    """BASELINE config 5 (per-trial aligned projection + decode): FusedPredictor on models fitted
    by the drop-in classes against the float64 CPU port of the reference path -- >= 99 % of 120
    held-out labels per decoder class, at batch size 1 (one trial per call) and in one batch."""
    from sklearn.pipeline import make_pipeline
    from cross_patient_speech_decoding_b200.alignment.AlignCCA import AlignCCA
    from cross_patient_speech_decoding_b200.alignment.AlignMCCA import AlignMCCA
    from cross_patient_speech_decoding_b200.decoders.cross_pt_decoders import (
        crossPtDecoder_mcca, crossPtDecoder_sepAlign)
    from cross_patient_speech_decoding_b200.decoders.fused_predict import FusedPredictor
    from cross_patient_speech_decoding_b200.decomposition.DimRedReshape import DimRedReshape
    from cross_patient_speech_decoding_b200.decomposition.PCA import PCA
    from cross_patient_speech_decoding_b200.folds import cv_splits
    from cross_patient_speech_decoding_b200.svm import LinearSVC
    from oracle import pipeline_port as port
    pts = _patients(3, n_trials=120)
    Xt, yt, yat = pts[0]
    np.random.seed(12)
    folds = cv_splits(yt, 4)
    for cls, kw, method, nc in [
            (crossPtDecoder_mcca, dict(aligner=AlignMCCA, n_comp=8, regs=0.5, pca_var=0.8), 'mcca', 8),
            (crossPtDecoder_sepAlign, dict(aligner=AlignCCA, n_comp=0.9), 'cca', 0.9)]:
        same = tot = 0
        for tr, te in folds:
            m = cls(pts[1:], make_pipeline(DimRedReshape(PCA, n_components=0.8), LinearSVC()), **kw)
            m.fit(Xt[tr], yt[tr], y_align=yat[tr])
            fp = FusedPredictor(m)
            ref, _ = port.run_fold(pts[0], pts[1:], tr, te, method=method, n_comp=nc)
            batch = fp.predict(Xt[te])
            one = np.concatenate([fp.predict(Xt[i:i + 1]) for i in te[:6]])
            assert np.array_equal(one, batch[:6])                # batch 1 == batched
            same += int((batch == ref).sum())
            tot += len(te)
        assert tot >= 100 and same / tot >= 0.99, (method, same, tot)
