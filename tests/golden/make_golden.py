"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference; MCCA through oracle/mcca_restated.py because mvlearn is absent) on the
seeded synthetic patients.  Run from the repo root in the build container:

    python tests/golden/make_golden.py [name ...]

The GPU box has no /root/reference, so the parity tests there compare against these files.
Inputs are NOT stored: they are regenerated from seeds by
cross_patient_speech_decoding_b200.synthetic, fold indices by folds.py under np.random.seed.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cross_patient_speech_decoding_b200 import synthetic  # noqa: E402
from cross_patient_speech_decoding_b200.folds import cv_splits  # noqa: E402
from oracle import run_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> config.  `patients`: kwargs list for synthetic.make_patient (ragged shapes allowed)
CONFIGS = {
    'cca_p2_5fold': dict(method='cca', n_comp=0.9, n_splits=5, seed=0,
                         patients=[dict(p=0), dict(p=1)]),
    'cca_p2_fixed30': dict(method='cca', n_comp=30, n_splits=5, seed=1,
                           patients=[dict(p=0), dict(p=1)]),
    'cca_p3_ragged': dict(method='cca', n_comp=0.9, n_splits=4, seed=2,
                          patients=[dict(p=0, n_trials=90, n_time=60, n_chan=48),
                                    dict(p=1, n_trials=110, n_time=60, n_chan=64),
                                    dict(p=2, n_trials=70, n_time=60, n_chan=33)]),
    'none_p3_ragged': dict(method='none', n_comp=0.9, n_splits=4, seed=3,
                           patients=[dict(p=0, n_trials=90, n_time=60, n_chan=48),
                                     dict(p=1, n_trials=110, n_time=60, n_chan=64),
                                     dict(p=2, n_trials=70, n_time=60, n_chan=33)]),
    'mcca_p3_ragged': dict(method='mcca', n_comp=12, regs=0.5, pca_var=0.8, n_splits=4, seed=4,
                           patients=[dict(p=0, n_trials=90, n_time=60, n_chan=48),
                                     dict(p=1, n_trials=110, n_time=60, n_chan=64),
                                     dict(p=2, n_trials=70, n_time=60, n_chan=33)]),
    'mcca_p8_20fold': dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, n_splits=20, seed=5,
                           patients=[dict(p=i) for i in range(8)], max_folds=6),
    # the scripts' literal decoder SVC(kernel='rbf', class_weight='balanced')
    'cca_p3_svc_rbf': dict(method='cca', n_comp=0.9, n_splits=4, seed=8, svm='svc_rbf',
                           patients=[dict(p=0, n_trials=90, n_time=60, n_chan=48, noise=0.6),
                                     dict(p=1, n_trials=110, n_time=60, n_chan=64, noise=0.6),
                                     dict(p=2, n_trials=70, n_time=60, n_chan=33, noise=0.6)]),
    'cca_p2_noisy': dict(method='cca', n_comp=0.9, n_splits=5, seed=6,
                         patients=[dict(p=0, noise=1.0), dict(p=1, noise=1.0)]),
}


def build_inputs(cfg):
    pts = [synthetic.make_patient(**kw) for kw in cfg['patients']]
    np.random.seed(cfg['seed'])
    folds = cv_splits(pts[0][1], cfg['n_splits'])
    if cfg.get('max_folds'):
        folds = folds[:cfg['max_folds']]
    return pts, folds


def generate(name):
    cfg = CONFIGS[name]
    pts, folds = build_inputs(cfg)
    t0 = time.time()
    res = run_reference.run_folds(pts[0], pts[1:], folds, method=cfg['method'],
                                  n_comp=cfg.get('n_comp'), regs=cfg.get('regs', 0.5),
                                  pca_var=cfg.get('pca_var', 0.8), svm=cfg.get('svm', 'primal'))
    dt = time.time() - t0
    nf = len(folds)
    out = dict(n_folds=nf, seconds_per_fold=dt / nf, k2=np.array(res['k2']),
               pool_shape=np.array(res['pool_shape']))
    for f in range(nf):
        out['train_%d' % f] = folds[f][0]
        out['test_%d' % f] = folds[f][1]
        out['y_pred_%d' % f] = res['y_pred'][f]
        out['y_true_%d' % f] = res['y_true'][f]
        out['svm_w_%d' % f] = res['svm_w'][f]
        if cfg['method'] == 'mcca':
            out['ranks_%d' % f] = np.array(res['ranks'][f])
            out['evals_mcca_%d' % f] = res['evals_mcca'][f]
            for v, (l, mu) in enumerate(zip(res['loadings'][f], res['means'][f])):
                out['loadings_%d_%d' % (f, v)] = l.astype(np.float32)
                out['means_%d_%d' % (f, v)] = mu.astype(np.float32)
        elif cfg['method'] == 'cca':
            out['d_a_%d' % f] = res['d_a'][f]
            for i in range(len(res['rho'][f])):
                out['rho_%d_%d' % (f, i)] = res['rho'][f][i]
                out['Ma_%d_%d' % (f, i)] = res['Ma'][f][i].astype(np.float32)
                out['Mb_%d_%d' % (f, i)] = res['Mb'][f][i].astype(np.float32)
        else:
            out['d_a_%d' % f] = res['d_a'][f]
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    acc = np.mean(np.concatenate(res['y_pred']) == np.concatenate(res['y_true']))
    print('%s: %d folds, %.2f s/fold, acc %.3f, k2 %s' % (name, nf, dt / nf, acc, res['k2']),
          flush=True)


if __name__ == '__main__':
    names = sys.argv[1:] or list(CONFIGS)
    for n in names:
        generate(n)
