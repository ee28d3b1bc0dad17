
import sys; sys.path.insert(0,"/root/repo")
import numpy as np
from cross_patient_speech_decoding_b200 import synthetic
from cross_patient_speech_decoding_b200.engine import CVEngine
from cross_patient_speech_decoding_b200.folds import cv_splits
from oracle import pipeline_port as port
for method, kw in [("mcca", dict(n_comp=6, regs=0.5, pca_var=0.8)), ("mcca", dict(n_comp=5, regs=0.1, pca_var=1)), ("jointpca", dict(n_comp=6))]:
    pts = [synthetic.make_patient(p, n_trials=n, n_time=30, n_chan=c) for p, n, c in ((0, 70, 24), (1, 85, 32))]
    np.random.seed(3); folds = cv_splits(pts[0][1], 4); folds[1] = (folds[1][0], folds[1][1][:3])
    res = CVEngine(pts[0], pts[1:], method=method, use_tensor_cores=True, **kw).run(folds)
    ag = tot = 0
    for f, (tr, te) in enumerate(folds):
        yp, k2 = port.run_fold(pts[0], pts[1:], tr, te, method=method, **kw)
        ag += int((res["y_pred"][f] == yp).sum()); tot += len(te)
    print(method, kw, ag, tot)
for shape in [dict(n=(50, 61), t=17, c=(20, 36), q=3), dict(n=(64, 40, 55), t=33, c=(32, 28, 128), q=32), dict(n=(48, 52), t=24, c=(30, 26), q=8), dict(n=(90, 75, 66, 58), t=16, c=(16, 24, 12, 40), q=5)]:
    pts = [synthetic.make_patient(p, n_trials=n, n_time=shape["t"], n_chan=c) for p, (n, c) in enumerate(zip(shape["n"], shape["c"]))]
    np.random.seed(11); folds = cv_splits(pts[0][1], 3)
    kw = dict(method="mcca", n_comp=shape["q"], regs=0.5, pca_var=1 if shape["q"] > 8 else 0.8, max_batch=2)
    a = CVEngine(pts[0], pts[1:], use_tensor_cores=True, **kw).run(folds)
    b = CVEngine(pts[0], pts[1:], use_tensor_cores=False, pool_solver="full", **kw).run(folds)
    same = sum(int((x == y).sum()) for x, y in zip(a["y_pred"], b["y_pred"])); tot = sum(len(x) for x in a["y_pred"])
    print(shape, same, tot, a["k2"] == b["k2"])
