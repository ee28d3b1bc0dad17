"""Label agreement of the fragile noisy CCA configuration under the different decoder-PCA solvers."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import make_golden
from cross_patient_speech_decoding_b200.engine import CVEngine
name = 'cca_p2_noisy'
cfg = make_golden.CONFIGS[name]
pts, folds = make_golden.build_inputs(cfg)
g = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
nf = int(g['n_folds'])
yr = np.concatenate([g['y_pred_%d' % f] for f in range(nf)])
for kw in (dict(), dict(pool_solver='full'), dict(pool_solver='topk'), dict(topk_gap_tol=1e9)):
    eng = CVEngine(pts[0], pts[1:], method=cfg['method'], n_comp=cfg.get('n_comp'), **kw)
    res = eng.run(folds[:nf])
    yp = np.concatenate(res['y_pred'])
    print(kw, 'agree %d/%d' % ((yp == yr).sum(), len(yr)), 'k2', res['k2'], eng.stats.get('topk_log'), flush=True)
