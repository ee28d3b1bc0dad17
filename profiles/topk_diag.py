"""Which top-k eigen-problems converge in how many rounds, which need the fp64 Cholesky-QR Gram,
which fall back to the full solver (and why): engine.stats['topk_log'] per method / latent size."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cross_patient_speech_decoding_b200.engine import CVEngine
from cross_patient_speech_decoding_b200.folds import cv_splits
pts = bench.make_data()
dev = [(torch.from_numpy(np.ascontiguousarray(X)).cuda(), y, ya) for X, y, ya in pts]
np.random.seed(1)
folds = cv_splits(pts[0][1], 20)
for method, d in (('jointpca', 30), ('jointpca', 60), ('mcca', 60), ('mcca', 45)):
    kw = dict(method=method, n_comp=d, use_tensor_cores=True, max_batch=20)
    if method == 'mcca': kw.update(regs=0.5, pca_var=0.8)
    eng = CVEngine(dev[0], dev[1:], **kw)
    r = eng.run(folds)
    print(method, d, 'k2', r['k2'][:4], [ {k: v for k, v in i.items()} for i in eng.stats['topk_log']], flush=True)
