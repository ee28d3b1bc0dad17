"""Run-to-run determinism of the engine on two configurations (bitwise comparison of the
decoder-PCA spectrum, the SVM weights and the predictions of two identical runs)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import make_golden
from cross_patient_speech_decoding_b200.engine import CVEngine
for name, kw in (('cca_p2_noisy', {}), ('cca_p2_5fold', {}), ('mcca_p8_20fold', dict(use_tensor_cores=True)),
                 ('mcca_p8_20fold', dict(use_tensor_cores=False, pool_solver='full'))):
    cfg = make_golden.CONFIGS[name]
    pts, folds = make_golden.build_inputs(cfg)
    outs = []
    for rep in range(2):
        eng = CVEngine(pts[0], pts[1:], method=cfg['method'], n_comp=cfg.get('n_comp'),
                       regs=cfg.get('regs', 0.5), pca_var=cfg.get('pca_var', 0.8), **kw)
        res = eng.run(folds, return_details=True)
        d = res['details'][0]
        k2 = res['k2']
        outs.append((np.concatenate(res['y_pred']), d['pool_evals'], d['W'], d, k2))
    (ya, ea, wa, da, k2a), (yb, eb, wb, db, k2b) = outs
    for key in ('loadings', 'evals_mcca', 'mu', 'Wt', 'G', 'rho', 'ev_t'):
        if key in da and da[key] is not None:
            x, y = np.asarray(da[key], dtype=float), np.asarray(db[key], dtype=float)
            print('   ', key, 'max rel diff %.2e' % (np.abs(x - y).max() / (np.abs(x).max() + 1e-300)))
    wd = max(np.abs(wa[f, :, :k2a[f] + 1] - wb[f, :, :k2b[f] + 1]).max() / np.abs(wa[f, :, :k2a[f] + 1]).max() for f in range(len(k2a)))
    wa, wb = np.array([1.0 + wd]), np.array([1.0])
    print(name, kw, 'labels equal', np.array_equal(ya, yb), '| pool evals max rel diff %.2e' %
          (np.abs(ea - eb).max() / np.abs(ea).max()), '| svm W max rel diff %.2e' %
          (np.abs(wa - wb).max() / np.abs(wa).max()))
