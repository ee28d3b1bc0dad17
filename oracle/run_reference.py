"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Runs the reference's own classes (imported from /root/reference through
oracle/reference_path.py) over a list of CV folds, the way the loop body of
scripts/aligned_decode_svm_ncv.py:344-442 does, with the pinned decoder
``make_pipeline(DimRedReshape(PCA, n_components=decoder_var),
CertifiedLinearSVC(dual=False, C, tol=1e-10, max_iter=100000))``  (see oracle/svm_exact.py for
why the primal liblinear solver, certified per class by its gradient and re-solved where it
stalled, is what defines the converged optimum).
Returns per-fold predictions plus the intermediate quantities the parity tests compare.
"""
import warnings

import numpy as np
from sklearn.decomposition import PCA
from sklearn.pipeline import make_pipeline
from sklearn.svm import LinearSVC

from oracle import reference_path


def make_decoder(ref, decoder_var=0.8, C=1.0, svm='primal'):
    if svm == 'primal':
        from oracle.svm_exact import oracle_linear_svc
        clf = oracle_linear_svc(C)     # liblinear primal, certified / re-solved per class
    elif svm == 'dual_default':
        clf = LinearSVC(dual=True, C=C, random_state=0)
    elif svm == 'svc_rbf':
        # the decoder the reference's scripts build themselves
        # (scripts/aligned_decode_svm_ncv.py:313-317)
        from sklearn.svm import SVC
        clf = SVC(kernel='rbf', class_weight='balanced', C=C)
    elif svm == 'svc_linear':                         # scripts/aligned_decode_svm.py:262
        from sklearn.svm import SVC
        clf = SVC(kernel='linear', C=C)
    else:
        raise ValueError(svm)
    return make_pipeline(ref.DimRedReshape(PCA, n_components=decoder_var), clf)


def run_folds(target, cross, folds, method='mcca', n_comp=None, regs=0.5, pca_var=0.8,
              decoder_var=0.8, C=1.0, svm='primal', details=True, tar_in_train=True):
    """target/cross: (X, y, y_align) tuples of float64 arrays; folds: (train, test) index pairs."""
    ref = reference_path.load()
    Xt, yt, yat = target
    out = dict(y_pred=[], y_true=[], k2=[], d_a=[], d_b=[], rho=[], ranks=[], evals_mcca=[],
               loadings=[], means=[], Ma=[], Mb=[], svm_w=[], pool_shape=[], W_joint=[])
    for tr, te in folds:
        clf = make_decoder(ref, decoder_var, C, svm)
        if method == 'mcca':
            m = ref.decoders.crossPtDecoder_mcca(cross, clf, ref.AlignMCCA,
                                                 n_comp=30 if n_comp is None else n_comp,
                                                 regs=regs, pca_var=pca_var,
                                                 tar_in_train=tar_in_train)
        elif method == 'cca':
            m = ref.decoders.crossPtDecoder_sepAlign(cross, clf, ref.AlignCCA,
                                                     n_comp=0.9 if n_comp is None else n_comp,
                                                     tar_in_train=tar_in_train)
        elif method == 'jointpca':
            import functools
            jp = functools.partial(ref.JointPCA, dim_red=functools.partial(PCA, svd_solver='full'))
            m = ref.decoders.crossPtDecoder_jointDimRed(cross, clf, jp,
                                                        n_comp=40 if n_comp is None else n_comp,
                                                        tar_in_train=tar_in_train)
        elif method == 'none':
            m = ref.decoders.crossPtDecoder_sepDimRed(cross, clf,
                                                      n_comp=0.9 if n_comp is None else n_comp,
                                                      tar_in_train=tar_in_train)
        else:
            raise ValueError(method)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            m.fit(Xt[tr], yt[tr], y_align=yat[tr]) if method != 'none' else m.fit(Xt[tr], yt[tr])
            yp = m.predict(Xt[te])
        out['y_pred'].append(np.asarray(yp))
        out['y_true'].append(np.asarray(yt[te]))
        pca = clf.steps[0][1].transformer
        svc = clf.steps[1][1]
        out['k2'].append(int(pca.n_components_))
        out.setdefault('svm_refit', []).append(len(getattr(svc, 'refit_', [])))
        out['pool_shape'].append((pca.n_samples_, pca.n_features_in_))
        if details:
            if hasattr(svc, 'dual_coef_'):
                out['svm_w'].append(np.asarray(svc.intercept_))
            else:
                out['svm_w'].append(np.hstack([svc.coef_, svc.intercept_[:, None]]))
            if method == 'mcca':
                mc = m.aligner.mcca
                out['ranks'].append(None if mc.signal_ranks is None else
                                    [int(r) for r in mc.signal_ranks])
                out['evals_mcca'].append(np.asarray(mc.evals_))
                out['loadings'].append([np.asarray(l) for l in mc.loadings_])
                out['means'].append([np.asarray(mu) for mu in mc.means_])
            elif method == 'cca':
                out['d_a'].append(int(m.tar_dr.n_components_))
                out['rho'].append([np.asarray(a.canon_corrs) for a in m.algns])
                out['Ma'].append([np.asarray(a.M_a) for a in m.algns])
                out['Mb'].append([np.asarray(a.M_b) for a in m.algns])
            elif method == 'jointpca':
                out['W_joint'].append([np.asarray(w) for w in m.joint_dr.transforms])
            else:
                out['d_a'].append(int(m.common_dim))
    return out
