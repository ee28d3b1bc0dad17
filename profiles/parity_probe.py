import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden'); sys.path.insert(0, 'tests')
import make_golden
from cross_patient_speech_decoding_b200.engine import CVEngine
for name in ['cca_p2_noisy', 'cca_p2_noisy', 'cca_p2_noisy']:
    cfg = make_golden.CONFIGS[name]
    pts, folds = make_golden.build_inputs(cfg)
    g = np.load('tests/golden/%s.npz' % name)
    eng = CVEngine(pts[0], pts[1:], method=cfg['method'], n_comp=cfg.get('n_comp'))
    res = eng.run(folds, return_details=True)
    nf = len(folds)
    yp = np.concatenate(res['y_pred']); yr = np.concatenate([g['y_pred_%d' % f] for f in range(nf)]); yt = np.concatenate([g['y_true_%d' % f] for f in range(nf)])
    det = res['details'][0]
    rels = []
    for f in range(nf):
        Ma, Mb = g['Ma_%d_0' % f].astype(float), g['Mb_%d_0' % f].astype(float)
        Gref = Mb @ np.linalg.pinv(Ma); db, da = Gref.shape
        rels.append(np.abs(det['G'][f, 0, :db, :da] - Gref).max() / np.abs(Gref).max())
    print(eng.stats.get('topk'), end=' ')
    print(name, 'agree %.4f (%d/%d)' % (np.mean(yp == yr), (yp == yr).sum(), len(yp)), 'acc ours %.4f ref %.4f' % (np.mean(yp == yt), np.mean(yr == yt)), 'k2', res['k2'], list(g['k2']), 'G rel err', ['%.1e' % r for r in rels])
