"""JointPCA configuration (BASELINE config 3): rounds of the two top-k solves, SVM iterations, folds/s."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cross_patient_speech_decoding_b200.engine import CVEngine  # noqa: E402
from cross_patient_speech_decoding_b200.folds import cv_splits  # noqa: E402

pts = bench.make_data()
dev = [(torch.from_numpy(np.ascontiguousarray(X)).cuda(), y, ya) for X, y, ya in pts]
d = int(sys.argv[1]) if len(sys.argv) > 1 else 30
folds = []
for it in range(4):
    np.random.seed(100 + it)
    folds += cv_splits(pts[0][1], 20)
eng = CVEngine(dev[0], dev[1:], method='jointpca', n_comp=d, use_tensor_cores=True, max_batch=148)
eng.run(folds)
eng.run(folds)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = eng.run(folds, return_details=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
si = np.concatenate([x['svm_info'].reshape(-1, 4) for x in res['details']])
print('d=%d: %.0f folds/s; svm newton mean %.1f max %d, cg mean %.0f max %d; k2 %s'
      % (d, len(folds) / dt, si[:, 0].mean(), si[:, 0].max(), si[:, 1].mean(), si[:, 1].max(), res['k2'][:4]))
for lg in eng.stats.get('topk_log', [])[-4:]:
    print('  topk', {k: v for k, v in lg.items()})
eng.profile = True
eng.run(folds)
print('  stages', {k: round(v, 1) for k, v in eng.collect_marks().items()})
