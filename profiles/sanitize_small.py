"""Tiny run of every engine path for compute-sanitizer (memcheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cross_patient_speech_decoding_b200 import synthetic
from cross_patient_speech_decoding_b200.engine import CVEngine
from cross_patient_speech_decoding_b200.folds import cv_splits
pts = [synthetic.make_patient(p, n_trials=n, n_time=30, n_chan=c) for p, n, c in ((0, 70, 24), (1, 85, 32), (2, 64, 28))]
np.random.seed(3)
folds = cv_splits(pts[0][1], 4)
for method, kw in (('mcca', dict(n_comp=6, regs=0.5, pca_var=0.8, use_tensor_cores=True)),
                   ('mcca', dict(n_comp=6, regs=0.5, pca_var=0.8, use_tensor_cores=False, pool_solver='full')),
                   ('jointpca', dict(n_comp=6, use_tensor_cores=True)),
                   ('cca', dict(n_comp=0.9)), ('none', dict(n_comp=0.9))):
    eng = CVEngine(pts[0], pts[1:], method=method, max_batch=2, **kw)
    res = eng.run(folds)
    print(method, kw.get('use_tensor_cores'), 'k2', res['k2'], 'ok')
