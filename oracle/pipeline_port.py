"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Self-contained CPU port (numpy / scipy / scikit-learn, float64) of the reference's
cross-validated align -> reduce -> decode fold, for machines where /root/reference does
not exist (the GPU box): it is the checker for ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` arm of bench.py.  Every function cites the
reference lines it restates; tests/test_oracle.py pins it against the golden outputs the
real reference produced (tests/golden/*.npz) and, when /root/reference is present, against
the reference classes run live.  MCCA goes through oracle/mcca_restated.py (PARITY
UNPINNED for that third-party dependency, see its header).
"""
from functools import reduce

import numpy as np
from sklearn.decomposition import PCA
from sklearn.svm import LinearSVC

from oracle.mcca_restated import MCCARestated


# ---------------------------------------------------------------- alignment_utils.py
def labels_as_str(lab):
    """alignment_utils.py:64-99 -- 2-D label rows become joined strings, 1-D become str."""
    lab = np.asarray(lab)
    if lab.ndim > 1:
        return np.array([''.join(str(v) for v in row) for row in lab])
    return lab.astype(str)


def condition_average(X, lab):
    """alignment_utils.py:42-61 -- mean over trials per class, classes in np.unique order."""
    classes = np.unique(lab)
    return np.stack([X[lab == c].mean(axis=0) for c in classes])


def shared_condition_averages(Xs, labs):
    """alignment_utils.py:12-39 -- class averages restricted to classes present everywhere."""
    labs = [labels_as_str(l) for l in labs]
    avgs = [condition_average(X, l) for X, l in zip(Xs, labs)]
    common = reduce(np.intersect1d, labs)
    return [a[np.isin(np.unique(l), common, assume_unique=True)] for a, l in zip(avgs, labs)]


# ---------------------------------------------------------------------- AlignMCCA.py
def signal_rank(X2d, var):
    """AlignMCCA.py:156-174 -- 0-based index of the first cumulative-variance value > var of
    the UNCENTRED singular spectrum (no +1, kept on purpose)."""
    s = np.linalg.svd(X2d, compute_uv=False) ** 2
    s = s / s.sum()
    return int(np.argmax(np.cumsum(s) > var))


def mcca_fit(Xs, labs, n_components, regs, pca_var):
    """AlignMCCA.py:140-154 (get_MCCA_transforms)."""
    views = [a.reshape(-1, a.shape[-1]) for a in shared_condition_averages(Xs, labs)]
    ranks = None
    if 0 < pca_var < 1:
        ranks = [min(n_components, signal_rank(X.reshape(-1, X.shape[-1]), pca_var)) for X in Xs]
    return MCCARestated(n_components=n_components, regs=regs, signal_ranks=ranks).fit(views)


def mcca_transform(model, X, view):
    """AlignMCCA.py:114-126."""
    out = model.transform_view(X.reshape(-1, X.shape[-1]), view)
    return out.reshape(X.shape[:-1] + (-1,))


# ----------------------------------------------------------------------- AlignCCA.py
def cca_directions(La, Lb):
    """AlignCCA.py:235-285 (CCA_align) on (samples x dims) inputs."""
    La = La - La.mean(axis=0)
    Lb = Lb - Lb.mean(axis=0)
    d = min(np.linalg.matrix_rank(La), np.linalg.matrix_rank(Lb))
    Qa, Ra = np.linalg.qr(La)
    Qb, Rb = np.linalg.qr(Lb)
    U, S, Vt = np.linalg.svd(Qa.T @ Qb)
    Ma = np.linalg.pinv(Ra) @ U[:, :d]
    Mb = np.linalg.pinv(Rb) @ Vt.T[:, :d]
    return Ma, Mb, np.clip(S[:d], 0.0, 1.0)


def cca_fit(Xa, Xb, ya, yb):
    """AlignCCA.py:43-61 with type='class' (AlignCCA.py:156-183): class averages of both
    patients over the classes they share, time folded into samples."""
    sa, sb = labels_as_str(ya), labels_as_str(yb)
    La, Lb = condition_average(Xa, sa), condition_average(Xb, sb)
    _, ia, ib = np.intersect1d(np.unique(sa), np.unique(sb), assume_unique=True,
                               return_indices=True)
    La, Lb = La[ia], Lb[ib]
    return cca_directions(La.reshape(-1, La.shape[-1]), Lb.reshape(-1, Lb.shape[-1]))


def trial_subselect(Xa, Xb, sa, sb):
    """AlignCCA.py:205-232 (shared_trial_subselect): per shared class, in np.intersect1d order,
    one np.random.permutation of patient A's trials of the class, then one of patient B's; the
    first min(count) of each are kept (numpy's GLOBAL RNG, as in the reference)."""
    La, Lb = [], []
    for c in np.intersect1d(sa, sb):
        ia = np.random.permutation(np.where(sa == c)[0])
        ib = np.random.permutation(np.where(sb == c)[0])
        m = min(len(ia), len(ib))
        La.append(Xa[ia[:m]])
        Lb.append(Xb[ib[:m]])
    return np.vstack(La), np.vstack(Lb)


def cca_fit_trial(Xa, Xb, ya, yb):
    """AlignCCA.fit with type='trial' (AlignCCA.py:43-61, 186-232)."""
    La, Lb = trial_subselect(Xa, Xb, labels_as_str(ya), labels_as_str(yb))
    return cca_directions(La.reshape(-1, La.shape[-1]), Lb.reshape(-1, Lb.shape[-1]))


# ----------------------------------------------------------------------- JointPCA.py
def joint_pca_fit(Xs, labs, n_components):
    """JointPCA.py:165-211 (get_joint_PCA_transforms): PCA of the channel-concatenated class
    averages, then per patient W_p = pinv(X_p) @ latent."""
    views = [a.reshape(-1, a.shape[-1]) for a in shared_condition_averages(Xs, labs)]
    latent = PCA(n_components=n_components, svd_solver='full').fit_transform(np.hstack(views))
    return [np.linalg.pinv(v) @ latent for v in views]


def joint_pca_transform(W, X):
    """JointPCA.py:132,149 -- no centring."""
    return (X.reshape(-1, X.shape[-1]) @ W).reshape(X.shape[:-1] + (-1,))


# -------------------------------------------------------------- cross_pt_decoders.py
def pool_mcca(Xtr, ytr, yal, cross, n_comp, regs, pca_var):
    """crossPtDecoder_mcca.preprocess_train (cross_pt_decoders.py:395-433)."""
    Xs = [Xtr] + [c[0] for c in cross]
    model = mcca_fit(Xs, [yal] + [c[2] for c in cross], n_comp, regs, pca_var)
    Z = [mcca_transform(model, X, i).reshape(X.shape[0], -1) for i, X in enumerate(Xs)]
    return np.vstack(Z), np.hstack([ytr] + [c[1] for c in cross]), model


def pool_cca(Xtr, ytr, yal, cross, n_comp):
    """crossPtDecoder_sepAlign.preprocess_train (cross_pt_decoders.py:211-270)."""
    pcs = [PCA(n_components=n_comp) for _ in cross]
    Xc = [p.fit_transform(c[0].reshape(-1, c[0].shape[-1])).reshape(c[0].shape[0],
                                                                    c[0].shape[1], -1)
          for p, c in zip(pcs, cross)]
    pt = PCA(n_components=n_comp)
    Xt = pt.fit_transform(Xtr.reshape(-1, Xtr.shape[-1])).reshape(Xtr.shape[0], Xtr.shape[1], -1)
    rows, info = [Xt.reshape(Xt.shape[0], -1)], []
    for Xb, c in zip(Xc, cross):
        Ma, Mb, rho = cca_fit(Xt, Xb, yal, c[2])
        rows.append((Xb @ Mb @ np.linalg.pinv(Ma)).reshape(Xb.shape[0], -1))   # AlignCCA.py:93
        info.append((Ma, Mb, rho))
    return np.vstack(rows), np.hstack([ytr] + [c[1] for c in cross]), pt, info


def pool_none(Xtr, ytr, cross, n_comp):
    """crossPtDecoder_sepDimRed.preprocess_train (cross_pt_decoders.py:117-163)."""
    Xc = [PCA(n_components=n_comp).fit_transform(c[0].reshape(-1, c[0].shape[-1])) for c in cross]
    pt = PCA(n_components=n_comp)
    Xt = pt.fit_transform(Xtr.reshape(-1, Xtr.shape[-1]))
    dim = min([Xt.shape[1]] + [x.shape[1] for x in Xc])
    rows = [Xt[:, :dim].reshape(Xtr.shape[0], -1)]
    rows += [x[:, :dim].reshape(c[0].shape[0], -1) for x, c in zip(Xc, cross)]
    return np.vstack(rows), np.hstack([ytr] + [c[1] for c in cross]), pt, dim


def run_fold(target, cross, train, test, method='mcca', n_comp=None, regs=0.5, pca_var=0.8,
             decoder_var=0.8, C=1.0, decoder='linear', class_weight=None, bag_random_state=None,
             n_estimators=10):
    """One unit of scripts/aligned_decode_svm_ncv.py:344-442 with the pinned decoder
    DimRedReshape(PCA, decoder_var) -> LinearSVC(dual=False) (DimRedReshape.py:36-65).
    Returns (y_pred, k2)."""
    Xt, yt, yat = target
    Xtr, Xte = Xt[train], Xt[test]
    if method == 'mcca':
        n_comp = 30 if n_comp is None else n_comp
        Xp, yp, model = pool_mcca(Xtr, yt[train], yat[train], cross, n_comp, regs, pca_var)
        Zte = mcca_transform(model, Xte, 0).reshape(len(test), -1)
    elif method == 'cca':
        n_comp = 0.9 if n_comp is None else n_comp
        Xp, yp, pt, _ = pool_cca(Xtr, yt[train], yat[train], cross, n_comp)
        Zte = pt.transform(Xte.reshape(-1, Xte.shape[-1])).reshape(len(test), -1)
    elif method == 'none':
        n_comp = 0.9 if n_comp is None else n_comp
        Xp, yp, pt, dim = pool_none(Xtr, yt[train], cross, n_comp)
        Zte = pt.transform(Xte.reshape(-1, Xte.shape[-1]))[:, :dim].reshape(len(test), -1)
    elif method == 'jointpca':
        # crossPtDecoder_jointDimRed (cross_pt_decoders.py:288-364) around JointPCA
        n_comp = 40 if n_comp is None else n_comp
        Xs = [Xtr] + [c[0] for c in cross]
        W = joint_pca_fit(Xs, [yat[train]] + [c[2] for c in cross], n_comp)
        Xp = np.vstack([joint_pca_transform(w, X).reshape(X.shape[0], -1) for w, X in zip(W, Xs)])
        yp = np.hstack([yt[train]] + [c[1] for c in cross])
        Zte = joint_pca_transform(W[0], Xte).reshape(len(test), -1)
    else:
        raise ValueError(method)
    pca = PCA(n_components=decoder_var).fit(Xp)
    if decoder == 'linear':
        from oracle.svm_exact import oracle_linear_svc
        svm = oracle_linear_svc(C)     # liblinear primal, certified / re-solved per class
    elif decoder.startswith('bag_'):
        # scripts/aligned_decode_svm.py:262-265: BaggingClassifier(estimator=SVC(kernel='linear'),
        # n_estimators=10); random_state=None there (seeds from numpy's global RNG at fit time)
        from sklearn.ensemble import BaggingClassifier
        from sklearn.svm import SVC
        svm = BaggingClassifier(estimator=SVC(kernel=decoder.split('_')[-1], C=C, class_weight=class_weight),
                                n_estimators=n_estimators, random_state=bag_random_state)
    else:
        # the scripts' literal decoder (scripts/aligned_decode_svm_ncv.py:313-317: rbf, balanced;
        # aligned_decode_svm.py:262: linear) -- sklearn's SVC is libsvm itself
        from sklearn.svm import SVC
        svm = SVC(kernel={'svc_rbf': 'rbf', 'svc_linear': 'linear'}[decoder], C=C,
                  class_weight=class_weight)
    svm.fit(pca.transform(Xp), yp)
    return svm.predict(pca.transform(Zte)), int(pca.n_components_)
