"""Batch routine for the pairwise-CCA (``crossPtDecoder_sepAlign`` + ``AlignCCA``) and the
un-aligned (``crossPtDecoder_sepDimRed``) decoders of the reference
(decoders/cross_pt_decoders.py:89-285), all folds of a batch at once.

Per fold: PCA of the target's train trials (time bins are samples); per cross patient a
fold-invariant PCA (done once in CVEngine._cross_pca); per (fold, cross patient) the CCA
of the class-averaged latents restricted to the classes both share (AlignCCA.py:156-183,
235-285), the map  G = M_b pinv(M_a)  composed with the cross patient's PCA basis; then
every trial is projected straight into the pooled matrix.
"""
import numpy as np
import torch

from . import _lib
from .device import addr, ptr
from .engine import _w_host, F32, I32, _ceil, ctypes_int_ptr


def _drain(g):
    try:
        while True:
            next(g)
    except StopIteration as e:
        return e.value


def batch_cca(eng, batch, want_details):
    return _drain(batch_cca_gen(eng, batch, want_details))


def batch_cca_gen(eng, batch, want_details):
    """Generator form (same protocol as CVEngine._batch_mcca_gen): yields 'sync' right before
    every blocking read-back and 'host' after a stretch of pure host packing, so that the lane
    scheduler of CVEngine.run can keep a second batch in flight."""
    ctx, T, P, Cm = eng.ctx, eng.T, eng.P, eng.Cmax
    B = len(batch)
    nv = P - 1
    tv = eng.views[0]
    aligned = eng.method == 'cca'
    launches0 = ctx.launches()
    n_padC = 128 if Cm <= 128 else _ceil(Cm, 128)

    # ------------------------------------------------------------- stage A: target PCA
    pk = eng.packA
    pk.reset()
    eng._o_zero = pk.o_zero = pk.add_ints([0])
    tabs = _drain(eng._target_tables(pk, batch))
    n_tr = [len(tb['tr']) for tb in tabs]
    n_te = [len(tb['te']) for tb in tabs]
    Kmax = max(len(tb['present']) for tb in tabs)
    pk.reserve_ints()
    mu_t = eng.ws('c_mu_t', (B, Cm))
    cov, gram_c = eng.scatter('c_cov', B, n_padC)
    if Cm < n_padC:
        cov.zero_()
    r_mu = np.zeros(B, dtype=_lib.COLSUM_DESC)
    r_cov = np.zeros(B, dtype=_lib.GRAM_TN_DESC)
    for f, tb in enumerate(tabs):
        sg = pk.iaddr(tb['o_tr'])
        r_mu[f] = (addr(tv.X), sg, addr(mu_t, f * Cm), n_tr[f], T, tv.C, tv.C,
                   1.0 / (n_tr[f] * T), 0)
        r_cov[f] = (addr(tv.X), addr(tv.X), sg, sg, addr(mu_t, f * Cm), addr(mu_t, f * Cm),
                    addr(cov, f * n_padC * n_padC), n_tr[f], T, tv.C, tv.C, tv.C, tv.C, n_padC, 1,
                    1.0 / (n_tr[f] * T - 1), 0)
    d_mu, d_cov = pk.add_descs(r_mu), pk.add_descs(r_cov)
    if aligned:
        cmT, r_cm = eng._class_means_target(pk, tabs, B, Kmax)
        d_cm = pk.add_descs(r_cm)
    pk.upload()
    eng.mark('align_scatter_eig')
    ctx.call('cpsd_colsum', pk.daddr(d_mu), B, tv.C)
    ctx.call(gram_c, pk.daddr(d_cov), B, tv.C, tv.C)
    if aligned:
        ctx.call('cpsd_class_mean', pk.daddr(d_cm), B, Kmax, T * tv.C)
    ev_t, evec_t = eng.eig_any(cov, n_padC, ptr(None), tv.C, B, 'ct')
    k_t = eng.ws('c_kt', (B,), I32)
    eng._select_pca_k(ev_t, n_padC, ptr(None), tv.C, k_t, 1, 0, B, eng.n_comp)
    yield 'sync'
    d_a = k_t.cpu().numpy().astype(np.int32)          # the one mid-batch sync: latent sizes
    d2h = d_a.nbytes
    d_b = eng.cross_k if nv else np.zeros(0, dtype=np.int32)
    if aligned:
        d_out = d_a.copy()                              # pooled latent width per fold
    else:
        cmin = int(d_b.min()) if nv else 1 << 30
        d_out = np.minimum(d_a, cmin).astype(np.int32)  # common_dim (cross_pt_decoders.py:146-149)
    dmax = _ceil(int(max([int(d_a.max())] + [int(x) for x in d_b])), 4)
    dq = _ceil(int(d_out.max()), 4)

    # ------------------------------------------------------------- stage B
    pk = eng.packB
    pk.reset()
    eng._o_zero = pk.o_zero = pk.add_ints([0])
    o_tr = [pk.add_ints(tb['tr'] * T) for tb in tabs]
    o_te = [pk.add_ints(tb['te'] * T) for tb in tabs]
    o_allseg = [pk.add_ints(np.arange(eng.views[v].N, dtype=np.int32) * T) for v in range(P)]
    cross_N = [eng.views[v].N for v in range(1, P)]
    n_pool = [(nt if eng.tar_in_train else 0) + sum(cross_N) for nt in n_tr]
    n_te_max = max(n_te)
    n_pad = _ceil(max(a + b for a, b in zip(n_pool, n_te)), 128)
    F = T * dq
    ypool = np.zeros((B, n_pad), dtype=np.int32)
    o_pooldst = np.zeros((B, P), dtype=np.int64)
    o_tedst = []
    for f, tb in enumerate(tabs):
        row, ys = 0, []
        if eng.tar_in_train:
            o_pooldst[f, 0] = pk.add_ints((row + np.arange(n_tr[f])) * T)
            ys.append(tv.y[tb['tr']])
            row += n_tr[f]
        for v in range(1, P):
            o_pooldst[f, v] = pk.add_ints((row + np.arange(eng.views[v].N)) * T)
            ys.append(eng.views[v].y)
            row += eng.views[v].N
        ypool[f, :row] = np.concatenate(ys)
        o_tedst.append(pk.add_ints((row + np.arange(n_te[f])) * T))
    o_ypool = pk.add_ints(ypool)
    o_npool = pk.add_ints(n_pool)
    o_nall = pk.add_ints([a + b for a, b in zip(n_pool, n_te)])
    o_nte = pk.add_ints(n_te)
    o_dt = pk.add_ints(d_a)
    o_cdim_t = pk.add_ints([tv.C] * B)
    o_cdim_x = pk.add_ints([eng.views[v].C for v in range(1, P)])
    # pairwise shared classes
    pairs = []
    if aligned:
        for f, tb in enumerate(tabs):
            slot_t = -np.ones(len(eng.vocab), dtype=np.int64)
            slot_t[tb['present']] = np.arange(len(tb['present']))
            for i in range(nv):
                sh = np.array(sorted(set(tb['present'].tolist()) & eng.cross_classes[i]),
                              dtype=np.int64)
                if len(sh) == 0:
                    raise ValueError('fold %d shares no alignment class with cross patient %d'
                                     % (f, i))
                pairs.append(dict(f=f, i=i, K=len(sh), sh=sh,
                                  o_a=pk.add_ints(slot_t[sh] * T),
                                  o_b=pk.add_ints(eng.cm_row[i + 1][sh] * T)))
        KTmax = max(p['K'] for p in pairs) * T
        o_segdst = pk.add_ints(np.arange(KTmax // T, dtype=np.int32) * T)
    pk.reserve_ints()

    # PCA bases (sklearn sign convention, zero-padded to dmax columns)
    Wt = eng.ws('c_Wt', (B, Cm, dmax))
    Wx = eng.ws('c_Wx', (max(nv, 1), Cm, dmax))
    npair = len(pairs)
    Zall = eng.ws('pool_Z', (B, n_pad, F))
    Zall.zero_()
    r_pp = []

    def proj_rec(X, src, dst, mu, W, ldw, f, nseg, C, q):
        return (addr(X), src, dst, mu, W, addr(Zall, f * n_pad * F), nseg, T, C, q, C, ldw, dq, 0)

    if aligned and nv:
        mA = eng.ws('c_mA', (npair, Cm))
        mB = eng.ws('c_mB', (npair, Cm))
        Lcat = eng.ws('c_L', (npair, KTmax, 2 * dmax))
        S = eng.ws('c_S', (npair, 2 * dmax, 2 * dmax))
        Ma = eng.ws('c_Ma', (npair, dmax, dmax))
        Mb = eng.ws('c_Mb', (npair, dmax, dmax))
        G = eng.ws('c_G', (npair, dmax, dmax))
        rho = eng.ws('c_rho', (npair, dmax))
        cinfo = eng.ws('c_info', (npair, 4), I32)
        Wc = eng.ws('c_Wc', (npair, Cm, dmax))
        r_m = np.zeros(2 * npair, dtype=_lib.COLSUM_DESC)
        r_pl = np.zeros(2 * npair, dtype=_lib.PROJ_DESC)
        r_s = np.zeros(npair, dtype=_lib.GRAM_TN_DESC)
        r_c = np.zeros(npair, dtype=_lib.CCA_DESC)
        r_w = np.zeros(npair, dtype=_lib.PROJ_DESC)
        for j, p in enumerate(pairs):
            f, i, K = p['f'], p['i'], p['K']
            xv = eng.views[i + 1]
            cma = addr(cmT, f * Kmax * T * tv.C)
            cmb = addr(eng.cm[i + 1])
            sa, sb = pk.iaddr(p['o_a']), pk.iaddr(p['o_b'])
            r_m[2 * j] = (cma, sa, addr(mA, j * Cm), K, T, tv.C, tv.C, 1.0 / (K * T), 0)
            r_m[2 * j + 1] = (cmb, sb, addr(mB, j * Cm), K, T, xv.C, xv.C, 1.0 / (K * T), 0)
            lbase = addr(Lcat, j * KTmax * 2 * dmax)
            r_pl[2 * j] = (cma, sa, pk.iaddr(o_segdst), addr(mA, j * Cm), addr(Wt, f * Cm * dmax),
                           lbase, K, T, tv.C, dmax, tv.C, dmax, 2 * dmax, 0)
            r_pl[2 * j + 1] = (cmb, sb, pk.iaddr(o_segdst), addr(mB, j * Cm),
                               addr(Wx, i * Cm * dmax), lbase + 4 * dmax, K, T, xv.C, dmax, xv.C,
                               dmax, 2 * dmax, 0)
            sbase = addr(S, j * 4 * dmax * dmax)
            r_s[j] = (lbase, lbase, pk.iaddr(eng._o_zero), pk.iaddr(eng._o_zero), 0, 0, sbase, 1,
                      K * T, 2 * dmax, 2 * dmax, 2 * dmax, 2 * dmax, 2 * dmax, 1, 1.0, 0)
            r_c[j] = (sbase, sbase + 4 * (dmax * 2 * dmax + dmax), sbase + 4 * dmax, 0, 0,
                      addr(Ma, j * dmax * dmax), addr(Mb, j * dmax * dmax),
                      addr(G, j * dmax * dmax), addr(rho, j * dmax), addr(cinfo, j * 4),
                      int(d_a[f]), int(d_b[i]), 2 * dmax, dmax, dmax, 0, 1e-10, 0)
            # Wc = W_b G  (C_b x dmax)
            r_w[j] = (addr(Wx, i * Cm * dmax), pk.iaddr(eng._o_zero), pk.iaddr(eng._o_zero), 0,
                      addr(G, j * dmax * dmax), addr(Wc, j * Cm * dmax), 1, xv.C, dmax, dmax,
                      dmax, dmax, dmax, 0)
        d_m, d_pl = pk.add_descs(r_m), pk.add_descs(r_pl)
        d_s, d_c, d_w = pk.add_descs(r_s), pk.add_descs(r_c), pk.add_descs(r_w)
    # pooled projection records
    for f, tb in enumerate(tabs):
        q = int(d_out[f])
        if eng.tar_in_train:
            r_pp.append(proj_rec(tv.X, pk.iaddr(o_tr[f]), pk.iaddr(o_pooldst[f, 0]),
                                 addr(mu_t, f * Cm), addr(Wt, f * Cm * dmax), dmax, f, n_tr[f],
                                 tv.C, q))
        for i in range(nv):
            xv = eng.views[i + 1]
            if aligned:
                W = addr(Wc, (f * nv + i) * Cm * dmax)
            else:
                W = addr(Wx, i * Cm * dmax)
            r_pp.append(proj_rec(xv.X, pk.iaddr(o_allseg[i + 1]), pk.iaddr(o_pooldst[f, i + 1]),
                                 addr(eng.cross_mu, i * Cm), W, dmax, f, xv.N, xv.C, q))
        r_pp.append(proj_rec(tv.X, pk.iaddr(o_te[f]), pk.iaddr(o_tedst[f]), addr(mu_t, f * Cm),
                             addr(Wt, f * Cm * dmax), dmax, f, n_te[f], tv.C, q))
    r_pp = np.array(r_pp, dtype=_lib.PROJ_DESC)
    d_pp = pk.add_descs(r_pp)
    r1, r2, pmu, Kall = eng._pooled_stage(pk, B, Zall, n_pad, F, n_pool, n_te, o_npool, o_nall,
                                          o_ypool, n_pad, n_te_max, want_details)
    d_p1, d_p2 = pk.add_descs(r1), pk.add_descs(r2)
    kcap = min(n_pad, F)
    St = eng.ws('pool_St', (B, kcap, n_pad))
    k2 = eng.ws('pool_k2', (B,), I32)
    W, info, r_svm = eng._svm_stage(pk, B, St, None, k2, kcap, n_pad, n_pool, n_te, o_ypool, n_pad,
                                    o_nte, n_te_max, ypool)
    d_svm = pk.add_descs(r_svm)
    pk.upload()
    yield 'host'

    # ------------------------------------------------------------- launches, stage B
    eng.mark('cca_solve')
    ctx.call('cpsd_pca_basis', ptr(evec_t), n_padC, n_padC * n_padC, ptr(k_t),
             ctypes_int_ptr(pk.iaddr(o_cdim_t)), 0, dmax, ptr(Wt), dmax, Cm, B)
    if nv:
        ctx.call('cpsd_pca_basis', ptr(eng.cross_evecs), eng.cross_npad,
                 eng.cross_npad * eng.cross_npad, ptr(eng.cross_k_dev),
                 ctypes_int_ptr(pk.iaddr(o_cdim_x)), 0, dmax, ptr(Wx), dmax, Cm, nv)
    if aligned and nv:
        ctx.call('cpsd_colsum', pk.daddr(d_m), 2 * npair, Cm)
        ctx.call('cpsd_proj_nn', pk.daddr(d_pl), 2 * npair, KTmax // T, T, dmax)
        ctx.call('cpsd_gram_tn', pk.daddr(d_s), npair, 2 * dmax, 2 * dmax)
        ctx.call('cpsd_cca_solve', pk.daddr(d_c), npair, dmax)
        ctx.call('cpsd_proj_nn', pk.daddr(d_w), npair, 1, Cm, dmax)
    eng.mark('project_pool')
    ctx.call('cpsd_proj_nn', pk.daddr(d_pp), len(r_pp),
             max(max(eng.views[v].N for v in range(P)), n_te_max), T, dq)
    evals, k2_, St_, Ste, V, sweeps, kcap = yield from eng._pooled_stage_run_gen(
        pk, d_p1, d_p2, B, Zall, pmu, Kall, n_pad, F, n_pool, n_te, o_npool, o_nall, o_ypool,
        n_te_max, want_details)
    ncls = len(eng.classes)
    eng.mark('svm')
    yhat = eng._decode(pk, d_svm, W, info, B, St_, Ste, k2, kcap, n_pad, o_ypool, n_pad, o_npool, o_nte,
                       n_te_max)
    eng.mark('end')
    yield 'sync'
    yh = yhat.cpu().numpy()
    k2h = k2.cpu().numpy()
    res = {'y_pred': [yh[f, :n_te[f]].copy() for f in range(B)], 'k2': k2h.tolist(),
           'h2d_bytes': eng.packA.h2d_bytes + pk.h2d_bytes,
           'd2h_bytes': yh.nbytes + k2h.nbytes + d2h}
    eng.stats['launches_last_batch'] = ctx.launches() - launches0
    if want_details:
        det = dict(d_a=d_a.copy(), d_b=np.array(d_b).copy(), d_out=d_out.copy(),
                   Wt=Wt.cpu().numpy(), mu_t=mu_t.cpu().numpy(), pool_evals=evals.cpu().numpy(),
                   svm_info=info.cpu().numpy(), W=_w_host(W), n_pool=list(n_pool),
                   ev_t=ev_t.cpu().numpy(),
                   bj_sweeps=None if sweeps is None else sweeps.cpu().numpy()[B:2 * B])
        if aligned and nv:
            det.update(rho=rho.view(B, nv, dmax).cpu().numpy(),
                       Ma=Ma.view(B, nv, dmax, dmax).cpu().numpy(),
                       Mb=Mb.view(B, nv, dmax, dmax).cpu().numpy(),
                       G=G.view(B, nv, dmax, dmax).cpu().numpy(),
                       cca_info=cinfo.view(B, nv, 4).cpu().numpy(),
                       Wx=Wx.cpu().numpy())
        res['details'] = det
    return res
