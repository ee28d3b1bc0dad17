import sys, os, json, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from cross_patient_speech_decoding_b200.engine import CVEngine
from cross_patient_speech_decoding_b200.folds import cv_splits
pts = bench.make_data()
dev = [(torch.from_numpy(np.ascontiguousarray(X)).cuda(), y, ya) for X, y, ya in pts]
def folds_for(y, n, iters, s0):
    out = []
    for it in range(iters):
        np.random.seed(s0 + it); out += cv_splits(y, n)
    return out
for tag, cross, nsp, iters in (('2 patients 5-fold', dev[1:2], 5, 16), ('8 patients 20-fold', dev[1:], 20, 4)):
  for nc in (0.9, 30):
    folds = folds_for(pts[0][1], nsp, iters, 100)
    eng = CVEngine(dev[0], cross, method='cca', n_comp=nc, use_tensor_cores=True, max_batch=148)
    eng.run(folds); eng.run(folds); torch.cuda.synchronize()     # (the second use moves the trial statistics to the all-trials eigenbasis, once)
    t0 = time.perf_counter(); res = eng.run(folds, return_details=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    eng.profile = True; eng.run(folds); st = eng.collect_marks()
    si = np.concatenate([d['svm_info'].reshape(-1, 4) for d in res['details']])
    ci = np.concatenate([d['cca_info'].reshape(-1, 4) for d in res['details']])
    print(tag, 'n_comp', nc, '%.0f folds/s' % (len(folds) / dt), {k: round(v, 2) for k, v in st.items()},
          'svm newton mean %.1f max %d cg mean %.0f max %d' % (si[:, 0].mean(), si[:, 0].max(), si[:, 1].mean(), si[:, 1].max()),
          'cca sweeps max', ci[:, 2].max(), flush=True)
