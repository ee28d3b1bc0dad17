"""Batch routine for the pairwise-CCA (``crossPtDecoder_sepAlign`` + ``AlignCCA``) and the
un-aligned (``crossPtDecoder_sepDimRed``) decoders of the reference
(decoders/cross_pt_decoders.py:89-285), all folds of a batch at once.

Per fold: PCA of the target's train trials (time bins are samples); per cross patient a
fold-invariant PCA (done once in CVEngine._cross_pca); per (fold, cross patient) the CCA
of the class-averaged latents restricted to the classes both share (AlignCCA.py:156-183,
235-285), the map  G = M_b pinv(M_a)  composed with the cross patient's PCA basis; then
every trial is projected straight into the pooled matrix.
"""
import ctypes

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
    # covariance of every fold's train trials from the per-trial Grams / column sums of the target
    # (CVEngine._target_trial_grams): rows [trial] ..., then [N] = minus the total -- a fold that
    # partitions the trials lists its held-out trials plus the total (sign -1), any other fold its
    # train trials (sign +1)
    tg = getattr(eng, 'tg', None)
    downdate = tg is not None and 'sums' in tg and n_padC == 128 and eng.J == 1
    rotated = downdate and bool(tg.get('rot'))
    if downdate:
        N0 = tv.N
        part = all(len(tb['tr']) + len(tb['te']) == N0 and
                   len(np.union1d(tb['tr'], tb['te'])) == N0 for tb in tabs)
        use_te = part and sum(n_te) <= sum(n_tr)
        lists = [np.concatenate([[N0], tb['te']]) if use_te else tb['tr'] for tb in tabs]
        o_lptr = pk.add_ints(np.concatenate([[0], np.cumsum([len(l) for l in lists])]))
        o_list = pk.add_ints(np.concatenate(lists))
        o_nrows = pk.add_ints(np.asarray(n_tr) * T)
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
    if downdate:
        sgn = -1.0 if use_te else 1.0
        lp, ll = ctypes_int_ptr(pk.iaddr(o_lptr)), ctypes_int_ptr(pk.iaddr(o_list))
        ssum = eng.ws('c_ssum', (B, 128), torch.float64)
        ctx.call('cpsd_sum_mats_f64', ptr(None), ptr(tg['trial']), 128 * 128, lp, ll, sgn, ptr(cov),
                 128 * 128, 128 * 128, B)
        ctx.call('cpsd_sum_mats_f64', ptr(None), ptr(tg['sums']), 128, lp, ll, sgn, ptr(ssum), 128, 128, B)
        smu = None
        if rotated:        # statistics live in the all-trials eigenbasis; the means do not
            smu = eng.ws('c_ssum0', (B, 128), torch.float64)
            ctx.call('cpsd_sum_mats_f64', ptr(None), ptr(tg['sums0']), 128, lp, ll, sgn, ptr(smu), 128, 128, B)
        ctx.call('cpsd_cov_from_sums', ptr(cov), 128, 128 * 128, ptr(ssum), 128,
                 ctypes_int_ptr(pk.iaddr(o_nrows)), tv.C, ptr(mu_t), Cm, ptr(smu), 1, B)
    else:
        ctx.call('cpsd_colsum', pk.daddr(d_mu), B, tv.C)
        ctx.call(gram_c, pk.daddr(d_cov), B, tv.C, tv.C)
    if aligned:
        ctx.call('cpsd_class_mean', pk.daddr(d_cm), B, Kmax, T * tv.C)
    if rotated:            # nearly diagonal already: start the accumulator from that basis
        ev_t, evec_t = eng.ws('ct_ev', (B, n_padC)), eng.ws('ct_evec', (B, n_padC, n_padC))
        eng.eig_warm(cov, ptr(None), tv.C, B, ev_t, evec_t, V0=tg['Qf'])
    else:
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
    # (all index tables and records below are built column-wise with numpy: no per-fold or
    # per-pair Python work beyond a few slice assignments)
    n_tr_a, n_te_a = np.asarray(n_tr, dtype=np.int64), np.asarray(n_te, dtype=np.int64)
    o_tr = pk.add_ints(np.concatenate([tb['tr'] for tb in tabs]) * T) + \
        np.concatenate([[0], np.cumsum(n_tr_a)])[:-1]
    o_te = pk.add_ints(np.concatenate([tb['te'] for tb in tabs]) * T) + \
        np.concatenate([[0], np.cumsum(n_te_a)])[:-1]
    Ns = np.array([eng.views[v].N for v in range(P)], dtype=np.int64)
    o_allseg = pk.add_ints(np.arange(int(Ns.max()), dtype=np.int32) * T)   # prefix serves every view
    sumN = int(Ns[1:].sum())
    n_pool_a = (n_tr_a if eng.tar_in_train else 0 * n_tr_a) + sumN
    n_pool = n_pool_a.tolist()
    n_te_max = int(n_te_a.max())
    n_pad = _ceil(int((n_pool_a + n_te_a).max()), 128)
    F = T * dq
    # destination rows inside the pooled matrix: slices of one table of row starts
    o_rows = pk.add_ints(np.arange(n_pad + n_te_max, dtype=np.int64) * T)
    row0 = n_tr_a if eng.tar_in_train else 0 * n_tr_a
    xoff = np.concatenate([[0], np.cumsum(Ns[1:])])[:-1]                 # start of each cross patient
    o_pooldst = np.zeros((B, P), dtype=np.int64)
    o_pooldst[:, 0] = o_rows
    o_pooldst[:, 1:] = o_rows + row0[:, None] + xoff[None, :]
    o_tedst = o_rows + n_pool_a
    ypool = np.zeros((B, n_pad), dtype=np.int32)
    ycross = np.concatenate([eng.views[v].y for v in range(1, P)]).astype(np.int32) if nv else \
        np.zeros(0, dtype=np.int32)
    for f, tb in enumerate(tabs):
        r0 = int(row0[f])
        if eng.tar_in_train:
            ypool[f, :r0] = tv.y[tb['tr']]
        ypool[f, r0:r0 + sumN] = ycross
    o_ypool = pk.add_ints(ypool)
    o_npool = pk.add_ints(n_pool)
    o_nall = pk.add_ints(n_pool_a + n_te_a)
    o_nte = pk.add_ints(n_te)
    o_dt = pk.add_ints(d_a)
    o_cdim_t = pk.add_ints([tv.C] * B)
    o_cdim_x = pk.add_ints([eng.views[v].C for v in range(1, P)])
    # pairwise shared classes: pair j = f * nv + i, classes ascending
    npair = 0
    if aligned and nv:
        V = len(eng.vocab)
        present2d = np.zeros((B, V), dtype=bool)
        for f, tb in enumerate(tabs):
            present2d[f, tb['present']] = True
        xm = getattr(eng, '_cross_mask', None)
        if xm is None:
            xm = np.zeros((nv, V), dtype=bool)
            rows = np.zeros((nv, V), dtype=np.int64)
            for i in range(nv):
                xm[i, sorted(eng.cross_classes[i])] = True
                rows[i] = np.asarray(eng.cm_row[i + 1], dtype=np.int64)
            eng._cross_mask, eng._cross_rows = xm, rows
        sh = (present2d[:, None, :] & xm[None, :, :]).reshape(B * nv, V)
        Kp = sh.sum(axis=1).astype(np.int64)
        if (Kp == 0).any():
            j = int(np.nonzero(Kp == 0)[0][0])
            raise ValueError('fold %d shares no alignment class with cross patient %d'
                             % (j // nv, j % nv))
        npair = B * nv
        pj, pc = np.nonzero(sh)                                        # pair, class (ascending)
        pf, pi = pj // nv, pj % nv
        slot_t = np.cumsum(present2d, axis=1) - 1                      # class -> row of the fold's means
        starts = np.concatenate([[0], np.cumsum(Kp)])[:-1]
        o_a = pk.add_ints(slot_t[pf, pc] * T) + starts
        o_b = pk.add_ints(eng._cross_rows[pi, pc] * T) + starts
        KTmax = int(Kp.max()) * T
        o_segdst = pk.add_ints(np.arange(KTmax // T, dtype=np.int32) * T)
    # PCA bases (sklearn sign convention, zero-padded to dmax columns)
    Wt = eng.ws('c_Wt', (B, Cm, dmax))
    Wx = eng.ws('c_Wx', (max(nv, 1), Cm, dmax))
    Zall = eng.ws('pool_Z', (B, n_pad, F))
    # pooled projection on the tensor cores (csrc/tc_proj.cu) when the shapes allow it: every
    # (fold, patient) gets its (channels x dq) map and mean, every trial a destination row
    tc_proj = eng._tc_proj_ready(dq)
    if tc_proj:
        Nmax = int(Ns.max())
        dst = -np.ones((B, P, Nmax), dtype=np.int32)
        for f, tb in enumerate(tabs):
            if eng.tar_in_train:
                dst[f, 0, tb['tr']] = np.arange(len(tb['tr']))
            dst[f, 0, tb['te']] = n_pool[f] + np.arange(len(tb['te']))
        for i in range(nv):
            dst[:, 1 + i, :Ns[1 + i]] = (row0[:, None] + xoff[i] + np.arange(Ns[1 + i])[None, :])
        o_dst = pk.add_ints(dst)
        slot = np.empty((B, P), dtype=np.int32)           # row of the mean vector in [mu_t ; cross_mu]
        slot[:, 0] = np.arange(B)
        slot[:, 1:] = B + np.arange(nv)[None, :]
        o_slot = pk.add_ints(slot)
        o_cdim_all = pk.add_ints(np.tile(np.array([eng.views[v].C for v in range(P)], dtype=np.int32), B))
    else:
        Zall.zero_()
    ib = pk.reserve_ints()
    xC = np.array([eng.views[v].C for v in range(1, P)], dtype=np.int64)
    xX = np.array([addr(eng.views[v].X) for v in range(1, P)], dtype=np.int64)
    fi = np.arange(B, dtype=np.int64)

    if aligned and nv:
        mA = eng.ws('c_mA', (npair, Cm))
        mB = eng.ws('c_mB', (npair, Cm))
        Lcat = eng.ws('c_L', (npair, KTmax, 2 * dmax))
        # scatter of [L_a | L_b] accumulated in fp64 and solved in fp64 (csrc/solve64.cu): the
        # Gram form squares the latents' condition number, fp32 moved the b->a map by up to 5e-3
        # on noisy configurations with ~100 latent dimensions
        S = eng.ws('c_S', (npair, 2 * dmax, 2 * dmax), torch.float64)
        cca_ws = eng.ws('c_cca_ws', (int(ctx.lib.cpsd_cca_solve_f64_ws_elems(npair, dmax)),),
                        torch.float64)
        Ma = eng.ws('c_Ma', (npair, dmax, dmax))
        Mb = eng.ws('c_Mb', (npair, dmax, dmax))
        G = eng.ws('c_G', (npair, dmax, dmax))
        rho = eng.ws('c_rho', (npair, dmax))
        cinfo = eng.ws('c_info', (npair, 4), I32)
        Wc = eng.ws('c_Wc', (npair, Cm, dmax))
        jj = np.arange(npair, dtype=np.int64)
        jf, ji = jj // nv, jj % nv
        cma = addr(cmT) + 4 * Kmax * T * tv.C * jf
        cmb = np.array([addr(eng.cm[i + 1]) for i in range(nv)], dtype=np.int64)[ji]
        sa, sb = ib + 4 * o_a, ib + 4 * o_b
        zero_a = pk.iaddr(eng._o_zero)
        segdst = pk.iaddr(o_segdst)
        inv_kt = (1.0 / (Kp * T)).astype(np.float32)
        r_m = np.zeros((npair, 2), dtype=_lib.COLSUM_DESC)
        r_m['A'][:, 0], r_m['A'][:, 1] = cma, cmb
        r_m['segA'][:, 0], r_m['segA'][:, 1] = sa, sb
        r_m['out'][:, 0], r_m['out'][:, 1] = addr(mA) + 4 * Cm * jj, addr(mB) + 4 * Cm * jj
        r_m['nseg'], r_m['seg_len'] = Kp[:, None], T
        r_m['p'][:, 0] = r_m['lda'][:, 0] = tv.C
        r_m['p'][:, 1] = r_m['lda'][:, 1] = xC[ji]
        r_m['alpha'] = inv_kt[:, None]
        r_m = r_m.ravel()
        lbase = addr(Lcat) + 4 * KTmax * 2 * dmax * jj
        r_pl = np.zeros((npair, 2), dtype=_lib.PROJ_DESC)
        r_pl['X'][:, 0], r_pl['X'][:, 1] = cma, cmb
        r_pl['seg_src'][:, 0], r_pl['seg_src'][:, 1] = sa, sb
        r_pl['seg_dst'] = segdst
        r_pl['mu'][:, 0], r_pl['mu'][:, 1] = addr(mA) + 4 * Cm * jj, addr(mB) + 4 * Cm * jj
        r_pl['W'][:, 0], r_pl['W'][:, 1] = addr(Wt) + 4 * Cm * dmax * jf, addr(Wx) + 4 * Cm * dmax * ji
        r_pl['Y'][:, 0], r_pl['Y'][:, 1] = lbase, lbase + 4 * dmax
        r_pl['nseg'], r_pl['seg_len'] = Kp[:, None], T
        r_pl['C'][:, 0] = r_pl['ldx'][:, 0] = tv.C
        r_pl['C'][:, 1] = r_pl['ldx'][:, 1] = xC[ji]
        r_pl['q'], r_pl['ldw'], r_pl['ldy'] = dmax, dmax, 2 * dmax
        r_pl = r_pl.ravel()
        sbase = addr(S) + 8 * 4 * dmax * dmax * jj
        r_s = np.zeros(npair, dtype=_lib.GRAM_TN_DESC)
        r_s['A'] = r_s['B'] = lbase
        r_s['segA'] = r_s['segB'] = zero_a
        r_s['out'], r_s['nseg'], r_s['seg_len'] = sbase, 1, Kp * T
        r_s['p'] = r_s['q'] = r_s['lda'] = r_s['ldb'] = r_s['ldo'] = 2 * dmax
        r_s['sym'], r_s['alpha'] = 1, 1.0
        r_c = np.zeros(npair, dtype=_lib.CCA_DESC)
        r_c['Saa'], r_c['Sbb'], r_c['Sab'] = sbase, sbase + 8 * (dmax * 2 * dmax + dmax), sbase + 8 * dmax
        r_c['Ma'], r_c['Mb'] = addr(Ma) + 4 * dmax * dmax * jj, addr(Mb) + 4 * dmax * dmax * jj
        r_c['G'], r_c['rho'] = addr(G) + 4 * dmax * dmax * jj, addr(rho) + 4 * dmax * jj
        r_c['info'] = addr(cinfo) + 16 * jj
        r_c['da'], r_c['db'] = d_a[jf], np.asarray(d_b)[ji]
        r_c['lds'], r_c['ldm'], r_c['ldg'], r_c['rank_tol'] = 2 * dmax, dmax, dmax, 1e-13
        # Wc = W_b G  (C_b x dmax)
        r_w = np.zeros(npair, dtype=_lib.PROJ_DESC)
        r_w['X'] = addr(Wx) + 4 * Cm * dmax * ji
        r_w['seg_src'] = r_w['seg_dst'] = zero_a
        r_w['W'], r_w['Y'] = addr(G) + 4 * dmax * dmax * jj, addr(Wc) + 4 * Cm * dmax * jj
        r_w['nseg'], r_w['seg_len'] = 1, xC[ji]
        r_w['C'] = r_w['q'] = r_w['ldx'] = r_w['ldw'] = r_w['ldy'] = dmax
        d_m, d_pl = pk.add_descs(r_m), pk.add_descs(r_pl)
        d_s, d_c, d_w = pk.add_descs(r_s), pk.add_descs(r_c), pk.add_descs(r_w)
    # pooled projection records: per fold [target train] + cross patients + target test
    nrec = (1 if eng.tar_in_train else 0) + nv + 1
    r_pp = np.zeros((B, nrec), dtype=_lib.PROJ_DESC)
    r_pp['Y'] = (addr(Zall) + 4 * n_pad * F * fi)[:, None]
    r_pp['seg_len'], r_pp['ldw'], r_pp['ldy'] = T, dmax, dq
    r_pp['q'] = np.asarray(d_out, dtype=np.int64)[:, None]
    c0 = 0
    if eng.tar_in_train:
        r_pp['X'][:, 0], r_pp['seg_src'][:, 0] = addr(tv.X), ib + 4 * o_tr
        r_pp['seg_dst'][:, 0] = ib + 4 * o_pooldst[:, 0]
        r_pp['mu'][:, 0], r_pp['W'][:, 0] = addr(mu_t) + 4 * Cm * fi, addr(Wt) + 4 * Cm * dmax * fi
        r_pp['nseg'][:, 0], r_pp['C'][:, 0], r_pp['ldx'][:, 0] = n_tr_a, tv.C, tv.C
        c0 = 1
    if nv:
        r_pp['X'][:, c0:c0 + nv] = xX[None, :]
        r_pp['seg_src'][:, c0:c0 + nv] = pk.iaddr(o_allseg)
        r_pp['seg_dst'][:, c0:c0 + nv] = ib + 4 * o_pooldst[:, 1:]
        r_pp['mu'][:, c0:c0 + nv] = (addr(eng.cross_mu) + 4 * Cm * np.arange(nv, dtype=np.int64))[None, :]
        if aligned:
            r_pp['W'][:, c0:c0 + nv] = addr(Wc) + 4 * Cm * dmax * (fi[:, None] * nv + np.arange(nv)[None, :])
        else:
            r_pp['W'][:, c0:c0 + nv] = (addr(Wx) + 4 * Cm * dmax * np.arange(nv, dtype=np.int64))[None, :]
        r_pp['nseg'][:, c0:c0 + nv] = Ns[None, 1:]
        r_pp['C'][:, c0:c0 + nv] = r_pp['ldx'][:, c0:c0 + nv] = xC[None, :]
    r_pp['X'][:, -1], r_pp['seg_src'][:, -1] = addr(tv.X), ib + 4 * o_te
    r_pp['seg_dst'][:, -1] = ib + 4 * o_tedst
    r_pp['mu'][:, -1], r_pp['W'][:, -1] = addr(mu_t) + 4 * Cm * fi, addr(Wt) + 4 * Cm * dmax * fi
    r_pp['nseg'][:, -1], r_pp['C'][:, -1], r_pp['ldx'][:, -1] = n_te_a, tv.C, tv.C
    r_pp = r_pp.ravel()
    d_pp = pk.add_descs(r_pp)
    r1, r2, pmu, Kall = eng._pooled_stage(pk, B, Zall, n_pad, F, n_pool, n_te, o_npool, o_nall,
                                          o_ypool, n_pad, n_te_max, want_details)
    d_p1, d_p2 = pk.add_descs(r1), pk.add_descs(r2)
    kcap = min(n_pad, F)
    St = eng.ws('pool_St', (B, kcap, n_pad))
    k2 = eng.ws('pool_k2', (B,), I32)
    W, info, r_svm = eng._svm_stage(pk, B, St, None, k2, kcap, n_pad, n_pool, n_te, o_ypool, n_pad,
                                    o_nte, n_te_max, ypool, batch=batch)
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
        ctx.call('cpsd_gram_tn_f64', pk.daddr(d_s), npair, 2 * dmax, 2 * dmax)
        ctx.call('cpsd_cca_solve_f64', pk.daddr(d_c), npair, dmax, ptr(cca_ws))
        ctx.call('cpsd_proj_nn', pk.daddr(d_w), npair, 1, Cm, dmax)
    eng.mark('project_pool')
    if tc_proj:
        # L[(f, v)] = target basis / aligned cross map, all with row stride dmax (first dq columns)
        Lall = eng.ws('c_Lall', (B, P, Cm, dmax))
        Lall[:, 0].copy_(Wt)
        if nv:
            if aligned:
                Lall[:, 1:].copy_(Wc.view(B, nv, Cm, dmax))
            else:
                Lall[:, 1:].copy_(Wx[:nv].unsqueeze(0).expand(B, nv, Cm, dmax))
        mu_all = eng.ws('c_muall', (B + max(nv, 1), Cm))
        mu_all[:B].copy_(mu_t)
        if nv:
            mu_all[B:B + nv].copy_(eng.cross_mu[:nv])
        tcp = eng._tc_proj_ws(B * P, dq)
        ctx.call('cpsd_proj_tc_prep', ptr(Lall), dmax, Cm * dmax, ptr(mu_all),
                 ctypes_int_ptr(pk.iaddr(o_slot)), Cm, ctypes_int_ptr(pk.iaddr(o_cdim_all)), dq, tcp['ltc'],
                 ptr(tcp['lthi']), ptr(tcp['ltlo']), ptr(tcp['mul']), B * P)
        ctx.call('cpsd_proj_tc', ptr(tcp['xmaps']), ptr(tcp['ltmaps']), P, B, T, dq, tcp['ltc'],
                 ctypes.c_void_p(tcp['ntr'].ctypes.data), ctypes.c_void_p(tcp['nch'].ctypes.data),
                 Nmax, ctypes_int_ptr(pk.iaddr(o_dst)), ptr(tcp['mul']), ptr(Zall), n_pad * F, tcp['sms'])
    else:
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
    eng._verify_spec()
    yh = yhat.cpu().numpy()
    k2h = k2.cpu().numpy()
    eng._check_decoder(info, B)
    if aligned and nv:
        bad = cinfo.view(npair, 4)[:, 1].cpu().numpy()
        if bad.any():
            j = int(np.nonzero(bad)[0][0])
            raise np.linalg.LinAlgError(
                'CCA: rank-deficient class-averaged latents (fold %d, cross patient %d); the '
                'reference truncates to matrix_rank (AlignCCA.py:263-265), which the Gram-form '
                'solver does not reproduce' % (j // nv, j % nv))
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
