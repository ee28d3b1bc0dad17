"""Generates tests/golden/jointpca_p3_ragged.npz by running the UNMODIFIED reference classes
``alignment.JointPCA.JointPCA`` and ``decoders.cross_pt_decoders.crossPtDecoder_jointDimRed``
(imported from /root/reference) on seeded synthetic patients:

    python tests/golden/make_golden_jointpca.py

Two variants of the joint PCA are stored: sklearn's default solver policy (randomized SVD
for this shape, seeded through numpy's global RNG) and the exact ``svd_solver='full'``.
Inputs are regenerated from seeds by make_golden.build_inputs (config mcca_p3_ragged).
"""
import functools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, '/root/reference/aligned_decoding')
import make_golden  # noqa: E402
from alignment.JointPCA import JointPCA  # noqa: E402  (reference)
from decoders.cross_pt_decoders import crossPtDecoder_jointDimRed  # noqa: E402  (reference)
from decomposition.DimRedReshape import DimRedReshape  # noqa: E402  (reference)
from sklearn.decomposition import PCA  # noqa: E402
from sklearn.pipeline import make_pipeline  # noqa: E402
from oracle.svm_exact import oracle_linear_svc  # noqa: E402

N_COMP = 12


def main():
    cfg = make_golden.CONFIGS['mcca_p3_ragged']
    pts, folds = make_golden.build_inputs(cfg)
    Xs, yal = [p[0] for p in pts], [p[2] for p in pts]
    out = dict(n_comp=N_COMP, n_folds=len(folds))
    full = functools.partial(PCA, svd_solver='full')
    for tag, dr in (('full', full), ('auto', PCA)):
        np.random.seed(11)
        jp = JointPCA(n_components=N_COMP, dim_red=dr)
        Z = jp.fit_transform(Xs, yal)
        for v in range(len(pts)):
            out['W_%s_%d' % (tag, v)] = jp.transforms[v]
            out['Z_%s_%d' % (tag, v)] = Z[v][:4]
        out['Zsingle_%s' % tag] = jp.transform(Xs[1][:3], idx=1)
    # decoder: joint dim-red of the train trials + pooled PCA(0.8) + linear SVM (exact optimum)
    Xt, yt, yat = pts[0]
    for f, (tr, te) in enumerate(folds):
        clf = make_pipeline(DimRedReshape(PCA, n_components=0.8),
                            oracle_linear_svc(1.0))
        model = crossPtDecoder_jointDimRed(pts[1:], clf, functools.partial(JointPCA, dim_red=full),
                                           n_comp=N_COMP)
        model.fit(Xt[tr], yt[tr], y_align=yat[tr])
        out['y_pred_%d' % f] = model.predict(Xt[te])
        out['y_true_%d' % f] = yt[te]
        out['k2_%d' % f] = clf.named_steps['dimredreshape'].transformer.n_components_
    np.savez_compressed(os.path.join(HERE, 'jointpca_p3_ragged.npz'), **out)
    acc = np.mean(np.concatenate([out['y_pred_%d' % f] == out['y_true_%d' % f]
                                  for f in range(len(folds))]))
    print('jointpca_p3_ragged: acc %.3f, k2 %s' % (acc, [int(out['k2_%d' % f]) for f in range(len(folds))]))
    d = np.abs(out['Z_full_0'] - out['Z_auto_0']).max() / np.abs(out['Z_full_0']).max()
    print('randomized vs full solver, relative difference of transformed data: %.2e' % d)


if __name__ == '__main__':
    main()
