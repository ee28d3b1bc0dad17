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

_RT = (144, 148, 151, 46, 151, 137, 141, 178)
_RC = (111, 111, 63, 149, 74, 144, 171, 201)
REAL_SHAPES = [dict(p=i, n_trials=n, n_chan=c, noise=(0.5 if i in (2, 5) else 0.15))
               for i, (n, c) in enumerate(zip(_RT, _RC))]
WIDE3 = [dict(p=0, n_trials=120, n_time=100, n_chan=128, noise=0.5),
         dict(p=1, n_trials=132, n_time=100, n_chan=120, noise=0.5),
         dict(p=2, n_trials=100, n_time=100, n_chan=128, noise=0.5)]

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
    # the headline workload: all 20 folds of 2 CV iterations (288 held-out labels); matrices
    # (loadings, means, SVM weights) are stored for the first `heavy` folds only
    'mcca_p8_20fold': dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, n_splits=20, seed=5,
                           patients=[dict(p=i) for i in range(8)], n_iter=2, heavy=3),
    # the scripts' literal decoder SVC(kernel='rbf', class_weight='balanced')
    'cca_p3_svc_rbf': dict(method='cca', n_comp=0.9, n_splits=4, seed=8, svm='svc_rbf',
                           patients=[dict(p=0, n_trials=90, n_time=60, n_chan=48, noise=0.6),
                                     dict(p=1, n_trials=110, n_time=60, n_chan=64, noise=0.6),
                                     dict(p=2, n_trials=70, n_time=60, n_chan=33, noise=0.6)]),
    'cca_p2_noisy': dict(method='cca', n_comp=0.9, n_splits=5, seed=6,
                         patients=[dict(p=0, noise=1.0), dict(p=1, noise=1.0)]),
    # the trial / channel counts of the reference's real recordings:
    # ragged trials, channel counts that are odd and reach 201 (> 128)
    'cca_real_shapes': dict(method='cca', n_comp=0.9, n_splits=5, seed=21, heavy=2,
                            patients=REAL_SHAPES),
    'cca_real_shapes_t7': dict(method='cca', n_comp=0.9, n_splits=5, seed=22, heavy=1, target=7,
                               patients=REAL_SHAPES),
    'mcca_real_shapes': dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, n_splits=5, seed=23,
                             heavy=2, patients=REAL_SHAPES),
    'mcca_real_shapes_t7': dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, n_splits=5, seed=24,
                                heavy=1, target=7, patients=REAL_SHAPES),
    # latent sizes of BASELINE config 3's sweep above the tensor-core projection's 32 columns
    'cca_p3_d60': dict(method='cca', n_comp=60, n_splits=4, seed=31, heavy=1, patients=WIDE3),
    'cca_p3_d100': dict(method='cca', n_comp=100, n_splits=4, seed=32, heavy=1, patients=WIDE3),
    'jointpca_p3_d60': dict(method='jointpca', n_comp=60, n_splits=4, seed=33, heavy=1, patients=WIDE3),
    'jointpca_p3_d100': dict(method='jointpca', n_comp=100, n_splits=4, seed=34, heavy=1,
                             patients=WIDE3),
    'mcca_p3_d60_full': dict(method='mcca', n_comp=60, regs=0.5, pca_var=1, n_splits=4, seed=35,
                             heavy=1, patients=WIDE3),
}


def build_inputs(cfg):
    pts = [synthetic.make_patient(**kw) for kw in cfg['patients']]
    t = cfg.get('target', 0)
    pts = [pts[t]] + pts[:t] + pts[t + 1:]
    np.random.seed(cfg['seed'])
    folds = []
    for _ in range(cfg.get('n_iter', 1)):      # one shuffled StratifiedKFold per CV iteration, as
        folds += cv_splits(pts[0][1], cfg['n_splits'])   # scripts/aligned_decode_svm_ncv.py:336-342
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
    heavy = cfg.get('heavy', nf)          # folds that keep their matrices
    out = dict(n_folds=nf, seconds_per_fold=dt / nf, k2=np.array(res['k2']),
               pool_shape=np.array(res['pool_shape']), heavy=heavy,
               svm_refit=np.array(res.get('svm_refit', [0] * nf)))   # classes liblinear left unconverged
    for f in range(nf):
        out['train_%d' % f] = folds[f][0].astype(np.int16)
        out['test_%d' % f] = folds[f][1].astype(np.int16)
        out['y_pred_%d' % f] = res['y_pred'][f].astype(np.int8)
        out['y_true_%d' % f] = res['y_true'][f].astype(np.int8)
        if f < heavy:
            out['svm_w_%d' % f] = res['svm_w'][f]
        if cfg['method'] == 'mcca':
            if res['ranks'][f] is not None:
                out['ranks_%d' % f] = np.array(res['ranks'][f])
            out['evals_mcca_%d' % f] = res['evals_mcca'][f]
            if f < heavy:
                for v, (l, mu) in enumerate(zip(res['loadings'][f], res['means'][f])):
                    out['loadings_%d_%d' % (f, v)] = l.astype(np.float32)
                    out['means_%d_%d' % (f, v)] = mu.astype(np.float32)
        elif cfg['method'] == 'cca':
            out['d_a_%d' % f] = res['d_a'][f]
            for i in range(len(res['rho'][f])):
                out['rho_%d_%d' % (f, i)] = res['rho'][f][i]
                if f < heavy:
                    out['Ma_%d_%d' % (f, i)] = res['Ma'][f][i].astype(np.float32)
                    out['Mb_%d_%d' % (f, i)] = res['Mb'][f][i].astype(np.float32)
        elif cfg['method'] == 'jointpca':
            if f < heavy:
                for v, w in enumerate(res['W_joint'][f]):
                    out['W_%d_%d' % (f, v)] = w.astype(np.float32)
        else:
            out['d_a_%d' % f] = res['d_a'][f]
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    acc = np.mean(np.concatenate(res['y_pred']) == np.concatenate(res['y_true']))
    print('%s: %d folds, %.2f s/fold, acc %.3f, k2 %s, classes re-solved after a liblinear stall: %s'
          % (name, nf, dt / nf, acc, res['k2'], res.get('svm_refit')), flush=True)


if __name__ == '__main__':
    names = sys.argv[1:] or list(CONFIGS)
    for n in names:
        generate(n)
