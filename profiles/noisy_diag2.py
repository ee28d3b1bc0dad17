"""Where the noisy CCA configuration (tests/golden/cca_p2_noisy.npz) loses label agreement:
stage-by-stage comparison of the engine with the CPU port (float64) on the same folds --
pooled matrix, decoder-PCA spectrum, SVM weights -- and the port's own decision margins of the
held-out trials (a flip on a trial whose top-2 margin is at rounding level is not a defect)."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import make_golden  # noqa: E402
from sklearn.decomposition import PCA  # noqa: E402
from sklearn.svm import LinearSVC  # noqa: E402
from cross_patient_speech_decoding_b200.engine import CVEngine  # noqa: E402
from oracle import pipeline_port as port  # noqa: E402
from oracle.svm_exact import oracle_linear_svc  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'cca_p2_noisy'
cfg = make_golden.CONFIGS[name]
pts, folds = make_golden.build_inputs(cfg)
g = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
nf = int(g['n_folds'])
eng = CVEngine(pts[0], pts[1:], method=cfg['method'], n_comp=cfg.get('n_comp'), max_batch=nf)
res = eng.run(folds[:nf], return_details=True)
det = res['details'][0]
dq = (int(max(det['d_out'])) + 3) // 4 * 4
n_pad = (max(det['n_pool'][f] + len(folds[f][1]) for f in range(nf)) + 127) // 128 * 128
Xt, yt, yat = pts[0]
T = Xt.shape[1]
Z = eng._ws['pool_Z'][:nf * n_pad * T * dq].view(nf, n_pad, T * dq)
for f, (tr, te) in enumerate(folds[:nf]):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        Xp, yp, pt, info = port.pool_cca(Xt[tr], yt[tr], yat[tr], pts[1:], cfg.get('n_comp'))
        Zte = pt.transform(Xt[te].reshape(-1, Xt.shape[-1])).reshape(len(te), -1)
        pca = PCA(n_components=0.8).fit(Xp)
        svm = oracle_linear_svc(1.0).fit(pca.transform(Xp), yp)
        dec = svm.decision_function(pca.transform(Zte))
    srt = np.sort(dec, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    ref_pred = svm.classes_[np.argmax(dec, axis=1)]
    n_pool = det['n_pool'][f]
    da = int(det['d_a'][f])
    Zf = Z[f].cpu().numpy().astype(np.float64).reshape(Z.shape[1], T, dq)[:, :, :da]
    mine = Zf[:n_pool].reshape(n_pool, -1)
    ref = Xp - Xp.mean(axis=0)
    # PCA directions carry a sign freedom per latent column: align signs column-wise
    mine3, ref3 = mine.reshape(n_pool, T, da), ref.reshape(n_pool, T, da)
    sg = np.sign(np.einsum('ntd,ntd->d', mine3, ref3))
    err = np.abs(mine3 * sg - ref3).max(axis=(0, 1)) / np.abs(ref3).max()
    flips = np.nonzero(res['y_pred'][f] != ref_pred)[0]
    print('fold %d: d_a %d k2 %d/%d  pooled-matrix rel err per latent column: max %.2e (col %d), median %.2e'
          % (f, da, res['k2'][f], pca.n_components_, err.max(), int(err.argmax()), np.median(err)))
    ev = det['pool_evals'][f][:pca.n_components_]
    print('   decoder-PCA eigenvalue rel err %.2e' % (np.abs(ev / (n_pool - 1) - pca.explained_variance_).max()
                                                     / pca.explained_variance_[0]))
    print('   port margins (top1 - top2): min %.3e, 5 smallest %s; decision scale %.2f'
          % (margin.min(), np.array2string(np.sort(margin)[:5], precision=4), np.abs(dec).mean()))
    # decoder stage: train / test PCA scores per component (sign aligned), decision values
    k2 = res['k2'][f]
    n_te = len(te)
    kcap = n_pad
    St = eng._ws['pool_St'][:nf * kcap * n_pad].view(nf, kcap, n_pad)[f, :k2, :n_pool].cpu().numpy().astype(np.float64)
    nte_max = max(len(t) for _, t in folds[:nf])
    Ste = eng._ws['pool_Ste'][:nf * kcap * nte_max].view(nf, kcap, nte_max)[f, :k2, :n_te].cpu().numpy().astype(np.float64)
    Rtr, Rte = pca.transform(Xp).T, pca.transform(Zte).T          # (k2, n)
    sg = np.sign(np.sum(St * Rtr, axis=1))[:, None]
    e_tr = np.abs(St * sg - Rtr).max(axis=1) / np.abs(Rtr).max(axis=1)
    e_te = np.abs(Ste * sg - Rte).max(axis=1) / np.abs(Rtr).max(axis=1)
    print('   train-score rel err per component: max %.2e (comp %d), median %.2e | test-score: max %.2e (comp %d), median %.2e'
          % (e_tr.max(), int(e_tr.argmax()), np.median(e_tr), e_te.max(), int(e_te.argmax()), np.median(e_te)))
    print('   worst 5 test comps', np.argsort(e_te)[-5:], np.sort(e_te)[-5:])
    W = det['W'][f][:, :k2 + 1]
    Wk = np.hstack([W[:, :k2], W[:, -1:]]) if W.shape[1] != k2 + 1 else W
    Wfull = det['W'][f]
    dec_ours = Wfull[:, :k2] @ Ste + Wfull[:, k2][:, None]
    dref = dec.T
    print('   decision values: max abs diff ours vs port %.3e' % np.abs(dec_ours * 1.0 - dref).max())
    svm2 = oracle_linear_svc(1.0).fit(St.T, yp)
    dec2 = svm2.decision_function(Ste.T).T
    print('   liblinear on OUR scores vs our SVM: max abs diff %.3e; vs port decisions %.3e'
          % (np.abs(dec2 - dec_ours).max(), np.abs(dec2 - dref).max()))
    for i in flips:
        print('   FLIP trial %d: ours %d, port %d, port margin %.3e (rank %d of %d)'
              % (i, res['y_pred'][f][i], ref_pred[i], margin[i], int((margin < margin[i]).sum()), len(te)))
    for i in range(len(info)):
        Ma, Mb, rho = info[i]
        Gref = Mb @ np.linalg.pinv(Ma)
        G = det['G'][f, i, :Gref.shape[0], :Gref.shape[1]]
        print('   pair %d: rho err %.2e, G rel err %.2e' % (i, np.abs(det['rho'][f, i, :len(rho)] - rho).max(),
                                                             np.abs(G - Gref).max() / np.abs(Gref).max()))
