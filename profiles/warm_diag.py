"""Stage-by-stage comparison of the warm-started (rotated statistics) CCA path with the cold one on
a golden configuration: target PCA spectrum / subspace, means, canonical correlations, pooled spectrum."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import make_golden  # noqa: E402
from cross_patient_speech_decoding_b200.engine import CVEngine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'cca_p2_noisy'
cfg = make_golden.CONFIGS[name]
pts, folds = make_golden.build_inputs(cfg)
g = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
nf = int(g['n_folds'])
folds = folds[:nf]
kw = dict(method=cfg['method'], n_comp=cfg.get('n_comp'), regs=cfg.get('regs', 0.5), pca_var=cfg.get('pca_var', 0.8))
cold = CVEngine(pts[0], pts[1:], **kw)
cold.warm_start = False
dc = cold.run(folds, return_details=True)['details'][0]
warm = CVEngine(pts[0], pts[1:], **kw)
for use in range(2):
    dw = warm.run(folds, return_details=True)['details'][0]
    print('use', use, 'rotated', bool((warm.tg or {}).get('rot')))
    print('  d_a', dw['d_a'], dc['d_a'])
    print('  ev_t rel diff %.2e' % (np.abs(dw['ev_t'] - dc['ev_t']).max() / dc['ev_t'].max()))
    print('  mu_t diff %.2e' % np.abs(dw['mu_t'] - dc['mu_t']).max())
    for f in range(len(folds)):
        d = int(dc['d_a'][f])
        Ww, Wc = dw['Wt'][f][:, :d].astype(float), dc['Wt'][f][:, :d].astype(float)
        sv = np.linalg.svd(np.linalg.qr(Ww)[0].T @ np.linalg.qr(Wc)[0], compute_uv=False)
        rho_ref = g['rho_%d_0' % f]
        k = len(rho_ref)
        print('  fold %d: PCA subspace sin(max angle) %.2e  orth err %.2e  rho-ref warm %.2e cold %.2e  gap at cut %.2e'
              % (f, np.sqrt(max(0, 1 - sv.min() ** 2)), np.abs(Ww.T @ Ww - np.eye(d)).max(),
                 np.abs(dw['rho'][f, 0, :k] - rho_ref).max(), np.abs(dc['rho'][f, 0, :k] - rho_ref).max(),
                 (dc['ev_t'][f, d - 1] - dc['ev_t'][f, d]) / dc['ev_t'][f, 0]))
    pe = np.abs(dw['pool_evals'] - dc['pool_evals'])
    i = np.unravel_index(pe.argmax(), pe.shape)
    print('  pool_evals max diff %.3e at %s (value %.3e, top %.3e)' % (pe.max(), i, dc['pool_evals'][i], dc['pool_evals'].max()))
