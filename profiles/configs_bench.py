"""Throughput of the BASELINE.json configurations other than the headline one:
  1  two patients, pairwise CCA (PCA 0.9 and fixed 30), 5-fold
  3  latent-size sweep 10..100 x {jointpca, cca, mcca}, 20-fold
  4  electrode subsampling (grid / Poisson-disk) x 8 targets x 20 folds, pairwise CCA
  5  per-call latency of transform + predict of a fitted config-2 model, batch 1 and 256
Data resident in HBM, wall clock around a synchronised run (CUDA work only between the syncs),
one warm-up run first.  Prints one JSON line per measurement."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cross_patient_speech_decoding_b200 as cp  # noqa: E402
from cross_patient_speech_decoding_b200 import synthetic  # noqa: E402
from cross_patient_speech_decoding_b200.engine import CVEngine  # noqa: E402
from cross_patient_speech_decoding_b200.folds import cv_splits  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--configs', default='1,2s,3,4,5')
ap.add_argument('--iters', type=int, default=5, help='CV iterations per measurement')
ap.add_argument('--cca-batch', type=int, default=148)
ap.add_argument('--dcd', type=int, default=0, help='dual-CD warm-start epochs of the linear SVM')
ap.add_argument('--dims', default='10,20,30,40,50,60,70,80,90,100')
args = ap.parse_args()
todo = set(args.configs.split(','))
pts = bench.make_data()
dev = [(torch.from_numpy(np.ascontiguousarray(X)).cuda(), y, ya) for X, y, ya in pts]


def folds_for(y, n_splits, iters, seed0):
    out = []
    for it in range(iters):
        np.random.seed(seed0 + it)
        out += cv_splits(y, n_splits)
    return out


def timed(tag, eng, folds, extra=None):
    eng.run(folds)          # warm-up with the same batch structure: workspaces of every lane, caches
    eng.run(folds)          # (second use: trial statistics move to the all-trials eigenbasis, once)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = eng.run(folds)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    acc = float(np.mean(np.concatenate([yp == eng.views[0].y[te] for yp, (_, te) in zip(res['y_pred'], folds)])))
    rec = dict(config=tag, folds=len(folds), folds_per_s=round(len(folds) / dt, 1),
               ms_per_fold=round(1e3 * dt / len(folds), 3), accuracy=round(acc, 4),
               k2=int(np.median(res['k2'])))
    rec.update(extra or {})
    print(json.dumps(rec), flush=True)
    return res


if '1' in todo:
    for nc in (0.9, 30):
        folds = folds_for(pts[0][1], 5, 4 * args.iters, 100)
        eng = CVEngine(dev[0], dev[1:2], method='cca', n_comp=nc, use_tensor_cores=True, max_batch=args.cca_batch)
        timed('1: 2 patients, CCA, n_comp=%s, 5-fold' % nc, eng, folds)

if '2s' in todo:
    for dec, cw in (('svc_rbf', 'balanced'), ('svc_linear', None), ('linear', None)):
        folds = folds_for(pts[0][1], 20, args.iters, 150)
        eng = CVEngine(dev[0], dev[1:], method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder=dec,
                       class_weight=cw, use_tensor_cores=True, max_batch=148)
        timed('2: 8 patients, MCCA, 20-fold, decoder=%s class_weight=%s' % (dec, cw), eng, folds)
        eng.profile = True
        eng.run(folds)
        print('   stages ms:', {k: round(v, 2) for k, v in eng.collect_marks().items()}, flush=True)

if '3' in todo:
    for d in [int(x) for x in args.dims.split(',')]:
        for method in ('jointpca', 'cca', 'mcca'):
            folds = folds_for(pts[0][1], 20, args.iters, 200)
            kw = dict(method=method, n_comp=d, use_tensor_cores=True, max_batch=148, dcd_epochs=args.dcd)
            if method == 'mcca':
                kw.update(regs=0.5, pca_var=0.8)
            if method == 'cca':
                kw.update(max_batch=args.cca_batch)
            eng = CVEngine(dev[0], dev[1:], **kw)
            try:
                timed('3: 8 patients, %s, d=%d, 20-fold' % (method, d), eng, folds)
                eng.profile = True               # stage timers: a separate single-lane run
                eng.run(folds)
                print('   stages ms:', {k: round(v, 2) for k, v in eng.collect_marks().items()}, flush=True)
            except ValueError as e:          # mvlearn raises too when n_components exceeds the summed ranks
                print(json.dumps(dict(config='3: 8 patients, %s, d=%d' % (method, d), error=str(e)[:80])), flush=True)

if '4' in todo:
    from cross_patient_speech_decoding_b200.processing_utils.grid_subsampling import sig_channels_in_windows
    from cross_patient_speech_decoding_b200.processing_utils.subsample_decode import subsample_decode
    chan_map = np.arange(1, 129).reshape(8, 16)
    sig = np.arange(1, 129)
    subs = [np.sort(np.asarray(s_).ravel()) for s_ in sig_channels_in_windows(chan_map, sig, (4, 8), (2, 4))]
    for rep in range(2):                      # rep 0 = warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_all = 0
        for tgt in range(2 if rep == 0 else 8):
            order = [tgt] + [p for p in range(8) if p != tgt]
            np.random.seed(300 + tgt)
            out = subsample_decode(pts[order[0]], [pts[p] for p in order[1:]], subs[:6], [subs] * 7,
                                   n_folds=20, method='cca', n_comp=0.9, use_tensor_cores=True,
                                   max_batch=148, depth=6)
            n_all += 20 * len(out['accs'])
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
    print(json.dumps(dict(config='4: 6 grid subsamples (32 of 128 channels) x 8 targets x 20 folds, CCA, '
                                 'subsample_decode (resident patients re-uploaded per target, 6 jobs in flight)',
                          folds=n_all, folds_per_s=round(n_all / t_all, 1), ms_per_fold=round(1e3 * t_all / n_all, 3))), flush=True)

if '5' in todo:
    from sklearn.pipeline import make_pipeline
    from cross_patient_speech_decoding_b200.alignment.AlignMCCA import AlignMCCA
    from cross_patient_speech_decoding_b200.decoders.cross_pt_decoders import crossPtDecoder_mcca
    from cross_patient_speech_decoding_b200.decomposition.DimRedReshape import DimRedReshape
    from cross_patient_speech_decoding_b200.decomposition.PCA import PCA
    from cross_patient_speech_decoding_b200.svm import LinearSVC
    Xt, yt, yat = pts[0]
    np.random.seed(0)
    tr, te = cv_splits(yt, 20)[0]
    m = crossPtDecoder_mcca(pts[1:], make_pipeline(DimRedReshape(PCA, n_components=0.8), LinearSVC()),
                            AlignMCCA, n_comp=30, regs=0.5, pca_var=0.8)
    t0 = time.perf_counter()
    m.fit(Xt[tr], yt[tr], y_align=yat[tr])
    fit_ms = 1e3 * (time.perf_counter() - t0)
    for nb in (1, 256):
        Xb = np.ascontiguousarray(np.concatenate([Xt] * 2)[:nb])
        for _ in range(5):
            m.predict(Xb)
        ts = []
        for _ in range(200 if nb == 1 else 50):
            t0 = time.perf_counter()
            yp = m.predict(Xb)
            ts.append(1e3 * (time.perf_counter() - t0))
        ts = np.sort(ts)
        from cross_patient_speech_decoding_b200.decoders.fused_predict import FusedPredictor
        fp = FusedPredictor(m)
        assert np.mean(fp.predict(Xb) == yp) >= 0.97
        tf = []
        for _ in range(300 if nb == 1 else 80):
            t0 = time.perf_counter()
            fp.predict(Xb)
            tf.append(1e3 * (time.perf_counter() - t0))
        print(json.dumps(dict(config='5: fitted config-2 model, FusedPredictor.predict latency, host float64 in -> labels out',
                              batch=nb, p50_ms=round(float(np.percentile(tf, 50)), 3),
                              p99_ms=round(float(np.percentile(tf, 99)), 3),
                              trials_per_s=round(nb / (np.median(tf) * 1e-3), 1))), flush=True)
        print(json.dumps(dict(config='5: fitted config-2 model, predict() latency, host float64 in -> labels out',
                              batch=nb, p50_ms=round(float(np.percentile(ts, 50)), 3),
                              p99_ms=round(float(np.percentile(ts, 99)), 3),
                              trials_per_s=round(nb / (np.median(ts) * 1e-3), 1), fit_ms=round(fit_ms, 1))), flush=True)
