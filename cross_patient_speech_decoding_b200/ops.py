"""numpy-in / numpy-out entry points to the individual kernels (host buffers, copies
included).  They back the sklearn-style classes (alignment/, decomposition/, svm.py) --
which are the batch-of-one case of the engine -- and the kernel-level parity tests.
Everything here runs on the GPU through libcpsd_b200.so; nothing is computed on the host
beyond packing index tables.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .device import Context, HostPack, addr, ptr

F32, I32, F64 = torch.float32, torch.int32, torch.float64


def _ctx(device=None):
    return Context.get(device)


def _ceil(a, b):
    return (a + b - 1) // b * b


def _ivoid(address):
    return ctypes.c_void_p(address)


def dgemm_batched(A, B, trans_a=False, alpha=1.0, beta=0.0, D=None, a_idx=None, b_idx=None,
                  device=None):
    """C[p] = alpha * op(A[a_idx[p]]) * B[b_idx[p]] + beta * D[p] in fp64 (host arrays in / out)."""
    ctx = _ctx(device)
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    nprob = len(a_idx) if a_idx is not None else (len(b_idx) if b_idx is not None else A.shape[0])
    m, k = (A.shape[2], A.shape[1]) if trans_a else (A.shape[1], A.shape[2])
    n = B.shape[2]
    assert B.shape[1] == k
    Ad, Bd = ctx.upload(A), ctx.upload(B)
    Dd = None if D is None else ctx.upload(np.ascontiguousarray(D, dtype=np.float64))
    ai = None if a_idx is None else ctx.upload(np.asarray(a_idx, dtype=np.int32))
    bi = None if b_idx is None else ctx.upload(np.asarray(b_idx, dtype=np.int32))
    C = ctx.empty((nprob, m, n), F64)
    ctx.call('cpsd_dgemm_batched', int(trans_a), m, n, k, float(alpha), ptr(Ad), A.shape[2],
             A.shape[1] * A.shape[2], ptr(ai), ptr(Bd), n, k * n, ptr(bi), float(beta), ptr(Dd), n, m * n,
             ptr(None), ptr(C), n, m * n, ptr(None), nprob)
    return C.cpu().numpy()


def eig_sym_warm(A, V0, v0_idx=None, n=None, sel=None, out_idx=None, n_out=None, device=None):
    """Warm-started fp64 tile solve (n <= 128): rotates A[p] into the basis V0[v0_idx[p]] (fp32
    eigenvector matrices of nearby problems, re-orthonormalised in fp64 on the device), solves from
    that starting accumulator and returns (evals, evecs, sweeps) like ``eig_sym(f64=True)``;
    ``v0_idx[p] < 0`` = cold solve of that problem."""
    ctx = _ctx(device)
    A = np.asarray(A, dtype=np.float64)
    nprob, nn, _ = A.shape
    assert nn <= 128
    ns = np.full(nprob, nn, dtype=np.int32) if n is None else np.asarray(n, dtype=np.int32)
    Ap = np.zeros((nprob, 128, 128))
    Ap[:, :nn, :nn] = A
    V0 = np.asarray(V0, dtype=np.float32)
    nb = V0.shape[0]
    V0p = np.zeros((nb, 128, 128), dtype=np.float32)
    V0p[:, :nn, :nn] = V0
    vi = np.zeros(nprob, dtype=np.int32) if v0_idx is None else np.asarray(v0_idx, dtype=np.int32)
    Ad, nd, V0d, vid = ctx.upload(Ap), ctx.upload(ns), ctx.upload(V0p), ctx.upload(vi)
    st = 128 * 128
    # fp64 re-orthonormalisation of the bases (one Newton-Schulz step)
    Qd, G, Q, Qf = (ctx.empty((nb, 128, 128), F64), ctx.empty((nb, 128, 128), F64),
                    ctx.empty((nb, 128, 128), F64), ctx.empty((nb, 128, 128)))
    ctx.call('cpsd_cast_f32_f64_idx', ptr(V0d), st, ptr(None), ptr(Qd), st, st, nb)
    ctx.call('cpsd_dgemm_batched', 1, 128, 128, 128, 1.0, ptr(Qd), 128, st, ptr(None), ptr(Qd), 128, st,
             ptr(None), 0.0, ptr(None), 128, 0, ptr(None), ptr(G), 128, st, ptr(None), nb)
    ctx.call('cpsd_dgemm_batched', 0, 128, 128, 128, -0.5, ptr(Qd), 128, st, ptr(None), ptr(G), 128, st,
             ptr(None), 1.5, ptr(Qd), 128, st, ptr(None), ptr(Q), 128, st, ptr(None), nb)
    ctx.call('cpsd_cast_f64_f32', ptr(Q), ptr(Qf), nb * st)
    warm = np.nonzero(vi >= 0)[0].astype(np.int32)
    if len(warm):
        wd, bd = ctx.upload(warm), ctx.upload(vi[warm])
        W = ctx.empty((len(warm), 128, 128), F64)
        ctx.call('cpsd_dgemm_batched', 0, 128, 128, 128, 1.0, ptr(Ad), 128, st, ptr(wd), ptr(Q), 128, st,
                 ptr(bd), 0.0, ptr(None), 128, 0, ptr(None), ptr(W), 128, st, ptr(None), len(warm))
        ctx.call('cpsd_dgemm_batched', 1, 128, 128, 128, 1.0, ptr(Q), 128, st, ptr(bd), ptr(W), 128, st,
                 ptr(None), 0.0, ptr(None), 128, 0, ptr(None), ptr(Ad), 128, st, ptr(wd), len(warm))
    sl = None if sel is None else ctx.upload(np.asarray(sel, dtype=np.int32))
    oi = None if out_idx is None else ctx.upload(np.asarray(out_idx, dtype=np.int32))
    nsel = nprob if sel is None else len(sel)
    n_out = nprob if n_out is None else n_out
    evals, evecs = ctx.zeros((n_out, 128)), ctx.zeros((n_out, 128, 128))
    sw = ctx.zeros((nprob,), I32)
    ctx.call('cpsd_eig_sym_small_f64_warm', ptr(Ad), 128, st, ptr(nd), 0, ptr(sl), nsel, ptr(oi),
             ptr(evals), 128, ptr(evecs), 128, st, ptr(Qf), 128, st, ptr(vid), 18, 1e-10, ptr(sw))
    return evals.cpu().numpy()[:, :nn], evecs.cpu().numpy()[:, :nn, :nn], sw.cpu().numpy()


def eig_sym(A, n=None, max_sweeps=15, tol=3e-7, device=None, return_sweeps=False, f64=False,
            tensor_cores=False):
    """Eigen-decomposition of a batch of symmetric matrices.  A: (nprob, n, n) or (n, n).
    ``f64`` (n <= 128): iterate the matrix in fp64, eigenvectors in fp32.
    Returns (evals descending (nprob, n), evecs (nprob, n, n) columns)."""
    ctx = _ctx(device)
    A = np.asarray(A, dtype=np.float64 if f64 else np.float32)
    single = A.ndim == 2
    if single:
        A = A[None]
    nprob, nn, _ = A.shape
    ns = np.full(nprob, nn, dtype=np.int32) if n is None else np.asarray(n, dtype=np.int32)
    n_pad = 128 if nn <= 128 else _ceil(nn, 128)
    if f64 and n_pad > 128:
        assert n_pad <= 256, 'fp64 eigen-solver: n <= 256'
        Ad, nd = ctx.upload(np.ascontiguousarray(A)), ctx.upload(ns)
        evals, evecs = ctx.empty((nprob, nn)), ctx.zeros((nprob, nn, nn))
        sw = ctx.zeros((nprob,), I32)
        ws = ctx.empty((int(ctx.lib.cpsd_eig_sym_f64_ws_elems(nprob, nn)),), F64)
        ctx.call('cpsd_eig_sym_f64', ptr(Ad), nn, nn * nn, ptr(nd), 0, nprob, ptr(evals), nn, ptr(evecs),
                 nn, nn * nn, 40, ptr(ws), nn, ptr(sw))
        ev, V, sweeps = evals.cpu().numpy(), evecs.cpu().numpy(), sw.cpu().numpy()
        if single:
            ev, V, sweeps = ev[0], V[0], sweeps[0]
        return (ev, V, sweeps) if return_sweeps else (ev, V)
    Ap = np.zeros((nprob, n_pad, n_pad), dtype=A.dtype)
    Ap[:, :nn, :nn] = A
    Ad = ctx.upload(Ap)
    nd = ctx.upload(ns)
    evals = ctx.empty((nprob, n_pad))
    sw = ctx.zeros((max(2 * nprob, 2),), I32)
    if n_pad <= 128:
        evecs = ctx.zeros((nprob, n_pad, n_pad))
        ctx.call('cpsd_eig_sym_small_f64' if f64 else 'cpsd_eig_sym_small', ptr(Ad), n_pad,
                 n_pad * n_pad, ptr(nd), 0, nprob, ptr(evals), n_pad, ptr(evecs), n_pad,
                 n_pad * n_pad, max_sweeps, 1e-10 if f64 else tol, ptr(sw))
        ev, V = evals.cpu().numpy(), evecs.cpu().numpy()
        sweeps = sw.cpu().numpy()[:nprob]
    else:
        nb = n_pad // 64
        sched = np.zeros((nb - 1) * (nb // 2) * 2, dtype=np.int32)
        _lib.check(ctx.lib.cpsd_bj_schedule(n_pad, sched.ctypes.data), 'bj_schedule')
        sd = ctx.upload(sched)
        R = ctx.empty((int(ctx.lib.cpsd_bj_rlog_elems(n_pad, nprob, max_sweeps)),))
        fw = ctx.zeros((18 * nprob,))
        perm = ctx.empty((nprob, n_pad), I32)
        RT = ctx.empty((nprob * (n_pad // 128) * 128 * 128,)) if tensor_cores else None
        ctx.call('cpsd_eig_sym_block', ptr(Ad), n_pad, n_pad * n_pad, n_pad, ptr(nd), 0, nprob,
                 ptr(sd), ptr(R), ptr(fw), ptr(sw), ptr(evals), ptr(perm), n_pad, max_sweeps, tol,
                 ptr(RT))
        evecs = ctx.zeros((nprob, n_pad, n_pad))
        ctx.call('cpsd_bj_eigvecs', ptr(R), n_pad, nprob, ptr(sd), ptr(sw), ptr(perm), n_pad,
                 ptr(None), n_pad, n_pad, ptr(evecs), n_pad, n_pad * n_pad, max_sweeps)
        ev, V = evals.cpu().numpy(), evecs.cpu().numpy()
        sweeps = sw.cpu().numpy()[nprob:2 * nprob]
        eig_sym.last_history = fw.cpu().numpy()[2 * nprob:].reshape(nprob, 16)
    ev, V = ev[:, :nn], V[:, :nn, :nn]
    if single:
        ev, V, sweeps = ev[0], V[0], sweeps[0]
    return (ev, V, sweeps) if return_sweeps else (ev, V)


def sgemm_batched(A, B, trans_a=False, alpha=1.0, device=None):
    """C[b] = alpha * op(A[b]) @ B[b] (fp32).  trans_a: A[b] is stored (K, M)."""
    ctx = _ctx(device)
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    nprob = A.shape[0]
    Kd, M = (A.shape[1], A.shape[2]) if trans_a else (A.shape[2], A.shape[1])
    N = B.shape[2]
    assert B.shape[1] == Kd
    Ad, Bd = ctx.upload(A), ctx.upload(B)
    C = ctx.empty((nprob, M, N))
    ctx.call('cpsd_sgemm_batched', int(bool(trans_a)), M, N, Kd, float(alpha), ptr(Ad), A.shape[2],
             A.shape[1] * A.shape[2], ptr(Bd), N, Kd * N, ptr(C), N, M * N, nprob)
    return C.cpu().numpy()


def chol_inv(S, device=None):
    """Inverse of the upper Cholesky factor: S[b] = R^T R -> (R^{-1} (nprob, m, m), status)."""
    ctx = _ctx(device)
    S = np.ascontiguousarray(S, dtype=np.float32)
    nprob, m, _ = S.shape
    Sd = ctx.upload(S)
    Rinv = ctx.empty((nprob, m, m))
    st = ctx.zeros((nprob,), I32)
    ctx.call('cpsd_chol_inv', ptr(Sd), m, m * m, m, ptr(Rinv), m, m * m, ptr(st), nprob)
    return Rinv.cpu().numpy(), st.cpu().numpy()


def lstsq_gram(X, Y, device=None):
    """Least-squares solution W = argmin |X W - Y| for a full-column-rank X (rows x C) through the
    normal equations in fp64: X^T X and X^T Y are accumulated in fp64 on the GPU and solved by a
    Cholesky factorisation (in shared memory up to 128 columns, in an L2-resident workspace above:
    patients with more than 128 electrodes).  Returns (W (C, q) float64, status)."""
    ctx = _ctx(device)
    X = np.ascontiguousarray(X, dtype=np.float32)
    Y = np.ascontiguousarray(Y, dtype=np.float32)
    rows, C = X.shape
    q = Y.shape[1]
    assert Y.shape[0] == rows
    Xd, Yd = ctx.upload(X), ctx.upload(Y)
    pk = HostPack(ctx)
    o = pk.add_ints([0])
    pk.reserve_ints()
    S = ctx.zeros((C, C), F64)
    Bm = ctx.zeros((C, q), F64)
    recs = np.zeros(2, dtype=_lib.GRAM_TN_DESC)
    recs[0] = (addr(Xd), addr(Xd), pk.iaddr(o), pk.iaddr(o), 0, 0, addr(S), 1, rows, C, C, C, C, C,
               1, 1.0, 0)
    recs[1] = (addr(Xd), addr(Yd), pk.iaddr(o), pk.iaddr(o), 0, 0, addr(Bm), 1, rows, C, q, C, q, q,
               0, 1.0, 0)
    d = pk.add_descs(recs)
    pk.upload()
    ctx.call('cpsd_gram_tn_f64', pk.daddr(d), 2, C, max(C, q))
    W = ctx.zeros((C, q))
    st = ctx.zeros((1,), I32)
    if C > 128:
        cw = ctx.empty((C * (C + 1),), F64)
        ctx.call('cpsd_chol_solve_f64_ws', ptr(S), C, C * C, C, ptr(Bm), q, C * q, q, ptr(W), q, C * q,
                 ptr(st), ptr(cw), 1)
    else:
        ctx.call('cpsd_chol_solve_f64', ptr(S), C, C * C, C, ptr(Bm), q, C * q, q, ptr(W), q, C * q,
                 ptr(st), 1)
    return W.cpu().numpy().astype(np.float64), int(st.cpu().numpy()[0])


def eig_topk(A, m=128, iters=8, rounds=1, n=None, max_sweeps=15, tol=3e-7, device=None,
             tensor_cores=False, tf32_iters=5, f64_gram=False):
    """Leading m eigen-pairs of symmetric PSD matrices A (nprob, n, n) by subspace iteration.
    Returns dict(evals (nprob, m), V (nprob, n, m), total, resid (nprob, m), status)."""
    ctx = _ctx(device)
    A = np.asarray(A, dtype=np.float32)
    nprob, nn, _ = A.shape
    n_pad = _ceil(nn, 128)
    ns = np.full(nprob, nn, dtype=np.int32) if n is None else np.asarray(n, dtype=np.int32)
    Ap = np.zeros((nprob, n_pad, n_pad), dtype=np.float32)
    Ap[:, :nn, :nn] = A
    Ad, nd = ctx.upload(Ap), ctx.upload(ns)
    ws = ctx.empty((int(ctx.lib.cpsd_eig_topk_ws_elems(n_pad, m, nprob)),))
    evals = ctx.empty((nprob, n_pad))
    tot, resid = ctx.empty((nprob,)), ctx.empty((nprob, m))
    st = ctx.zeros((nprob,), I32)
    if tensor_cores:
        tcw = ctx.empty((int(ctx.lib.cpsd_topk_tc_ws_elems(n_pad, nprob)),))
        nb = int(ctx.lib.cpsd_topk_tc_map_bytes(nprob))
        maps = ctx.empty((nb + 64,), torch.uint8)
        mp = (maps.data_ptr() + 63) & ~63
        stage = torch.empty((nb + 64,), dtype=torch.uint8).pin_memory()
        ctx.call('cpsd_topk_tc_encode', ptr(Ad), n_pad, n_pad * n_pad, n_pad, nprob, ptr(tcw),
                 ctypes.c_void_p(mp), ctypes.c_void_p(stage.data_ptr()))
    for r in range(rounds):
        if tensor_cores:
            ctx.call('cpsd_eig_sym_topk_tc', ptr(Ad), n_pad, n_pad * n_pad, n_pad, ptr(nd), 0, nprob,
                     m, iters, 1 if r == 0 else 0, ptr(ws), ptr(evals), n_pad, ptr(tot), ptr(resid),
                     ptr(st), max_sweeps, tol, ptr(tcw), ctypes.c_void_p(mp), tf32_iters, int(f64_gram))
        else:
            ctx.call('cpsd_eig_sym_topk', ptr(Ad), n_pad, n_pad * n_pad, n_pad, ptr(nd), 0, nprob,
                     m, iters, 1 if r == 0 else 0, ptr(ws), ptr(evals), n_pad, ptr(tot),
                     ptr(resid), ptr(st), max_sweeps, tol, int(f64_gram))
    voff = int(ctx.lib.cpsd_eig_topk_voff(n_pad, m, nprob))
    V = ws[voff:voff + nprob * 2 * n_pad * m].view(nprob, 2 * n_pad, m)[:, :nn, :]
    return dict(evals=evals.cpu().numpy()[:, :m], V=V.cpu().numpy(), total=tot.cpu().numpy(),
                resid=resid.cpu().numpy(), status=st.cpu().numpy())


def select_k(evals, thr, mode, n=None, kmin=0, kmax=1 << 30, device=None, total=None):
    """Component count from descending spectra; ``total``: per-problem total variance when only
    the leading eigenvalues are given."""
    ctx = _ctx(device)
    ev = np.atleast_2d(np.asarray(evals, dtype=np.float32))
    nprob, ld = ev.shape
    ns = np.full(nprob, ld, dtype=np.int32) if n is None else np.asarray(n, dtype=np.int32)
    evd, nd = ctx.upload(ev), ctx.upload(ns)
    k = ctx.empty((nprob,), I32)
    if total is not None:
        td = ctx.upload(np.asarray(total, dtype=np.float32))
        ctx.call('cpsd_select_k_total', ptr(evd), ld, ptr(nd), 0, ptr(td), float(thr), int(mode),
                 int(kmin), int(kmax), ptr(k), 1, nprob)
    else:
        ctx.call('cpsd_select_k', ptr(evd), ld, ptr(nd), 0, float(thr), int(mode), int(kmin),
                 int(kmax), ptr(k), 1, nprob)
    return k.cpu().numpy()


def gram_tn(A, B=None, seg_rows=None, seg_len=None, muA=None, muB=None, alpha=1.0, sym=None,
            device=None, f64=False):
    """out = alpha * sum over the selected rows of (a - muA)^T (b - muB).
    A: (rows, p); B: (rows, q) or None (= A).  seg_rows: first row of each segment."""
    ctx = _ctx(device)
    A = np.ascontiguousarray(A, dtype=np.float32)
    same = B is None
    Bm = A if same else np.ascontiguousarray(B, dtype=np.float32)
    if seg_rows is None:
        seg_rows, seg_len = np.zeros(1, dtype=np.int32), A.shape[0]
    p, q = A.shape[1], Bm.shape[1]
    Ad = ctx.upload(A)
    Bd = Ad if same else ctx.upload(Bm)
    pk = HostPack(ctx)
    o = pk.add_ints(seg_rows)
    pk.reserve_ints()
    ma = ctx.upload(np.asarray(muA, dtype=np.float32)) if muA is not None else None
    mb = ma if (same and muB is None) else (ctx.upload(np.asarray(muB, dtype=np.float32))
                                            if muB is not None else None)
    out = ctx.zeros((p, q), F64 if f64 else F32)
    if sym is None:
        sym = same and (muB is None)
    rec = np.zeros(1, dtype=_lib.GRAM_TN_DESC)
    rec[0] = (addr(Ad), addr(Bd), pk.iaddr(o), pk.iaddr(o), addr(ma), addr(mb), addr(out),
              len(seg_rows), int(seg_len), p, q, p, q, q, int(bool(sym)), float(alpha), 0)
    d = pk.add_descs(rec)
    pk.upload()
    ctx.call('cpsd_gram_tn_f64' if f64 else 'cpsd_gram_tn', pk.daddr(d), 1, p, q)
    return out.cpu().numpy()


def colmean(A, seg_rows=None, seg_len=None, device=None):
    ctx = _ctx(device)
    A = np.ascontiguousarray(A, dtype=np.float32)
    if seg_rows is None:
        seg_rows, seg_len = np.zeros(1, dtype=np.int32), A.shape[0]
    Ad = ctx.upload(A)
    pk = HostPack(ctx)
    o = pk.add_ints(seg_rows)
    pk.reserve_ints()
    out = ctx.zeros((A.shape[1],))
    rec = np.zeros(1, dtype=_lib.COLSUM_DESC)
    rec[0] = (addr(Ad), pk.iaddr(o), addr(out), len(seg_rows), int(seg_len), A.shape[1],
              A.shape[1], 1.0 / (len(seg_rows) * int(seg_len)), 0)
    d = pk.add_descs(rec)
    pk.upload()
    ctx.call('cpsd_colsum', pk.daddr(d), 1, A.shape[1])
    return out.cpu().numpy()


def gram_nt(A, B=None, alpha=1.0, device=None, tensor_cores=False):
    """out = alpha * A B^T  (A: (m, k), B: (n, k))."""
    ctx = _ctx(device)
    A = np.ascontiguousarray(A, dtype=np.float32)
    same = B is None
    Ad = ctx.upload(A)
    Bd = Ad if same else ctx.upload(np.ascontiguousarray(B, dtype=np.float32))
    m, k = A.shape
    n = m if same else B.shape[0]
    out = ctx.zeros((m, n))
    rec = np.zeros(1, dtype=_lib.GRAM_NT_DESC)
    rec[0] = (addr(Ad), addr(Bd), addr(out), m, n, k, k, k, n, int(same), float(alpha))
    pk = HostPack(ctx)
    pk.reserve_ints()
    d = pk.add_descs(rec)
    pk.upload()
    if tensor_cores:
        nbytes = int(ctx.lib.cpsd_gram_nt_tc_ws_bytes(1))
        split = ctx.empty((2 * m * k,))
        maps = ctx.empty((nbytes + 64,), torch.uint8)
        stage = torch.empty((nbytes + 64,), dtype=torch.uint8).pin_memory()
        a = (maps.data_ptr() + 63) & ~63
        ctx.call('cpsd_gram_nt_tc', ctypes.c_void_p(rec.ctypes.data), 1, m, n, ptr(split),
                 split.numel(), ctypes.c_void_p(a), ctypes.c_void_p(stage.data_ptr()))
        torch.cuda.synchronize()
    else:
        ctx.call('cpsd_gram_nt', pk.daddr(d), 1, m, n)
    return out.cpu().numpy()


def project(X, W, mu=None, device=None):
    """(X - mu) @ W for X (..., C), W (C, q), q <= 128."""
    ctx = _ctx(device)
    X = np.asarray(X, dtype=np.float32)
    lead, C = X.shape[:-1], X.shape[-1]
    X2 = np.ascontiguousarray(X.reshape(-1, C))
    W = np.ascontiguousarray(W, dtype=np.float32)
    q = W.shape[1]
    outs = []
    Xd, mud = ctx.upload(X2), (ctx.upload(np.asarray(mu, dtype=np.float32)) if mu is not None
                               else None)
    rows = X2.shape[0]
    for j0 in range(0, q, 128):
        Wd = ctx.upload(np.ascontiguousarray(W[:, j0:j0 + 128]))
        qq = Wd.shape[1]
        Y = ctx.empty((rows, qq))
        pk = HostPack(ctx)
        o = pk.add_ints([0])
        pk.reserve_ints()
        rec = np.zeros(1, dtype=_lib.PROJ_DESC)
        rec[0] = (addr(Xd), pk.iaddr(o), pk.iaddr(o), addr(mud), addr(Wd), addr(Y), 1, rows, C, qq,
                  C, qq, qq, 0)
        d = pk.add_descs(rec)
        pk.upload()
        ctx.call('cpsd_proj_nn', pk.daddr(d), 1, 1, rows, qq)
        outs.append(Y.cpu().numpy())
    Y = outs[0] if len(outs) == 1 else np.concatenate(outs, axis=1)
    return Y.reshape(lead + (q,))


def project_pool_tc(Xs, L, mu, dst, n_rows, device=None):
    """The tensor-core pooled projection on its own (csrc/tc_proj.cu, ``k_proj_tc``):
    Z[f][dst[f, v, trial]][t][:] = (X_v[trial][t][:] - mu[f, v]) @ L[f, v].
    Xs: list of P arrays (N_v, T, C_v) with C_v <= 256 (any parity); L: (B, P, Cmax, Q) with
    Q <= 128; mu: (B, P, Cmax) or None; dst: (B, P, Nmax) int destination trial rows (-1 = skip).
    Returns Z (B, n_rows, T * Q) float32 (rows no trial maps to stay zero)."""
    ctx = _ctx(device)
    P = len(Xs)
    L = np.ascontiguousarray(L, dtype=np.float32)
    B, P2, Cm, Q = L.shape
    assert P2 == P
    T = Xs[0].shape[1]
    Nmax = int(dst.shape[2])
    maps_h = torch.zeros((2 * P, 128), dtype=torch.uint8).pin_memory()
    keep = []
    for i, X in enumerate(Xs):
        X = np.ascontiguousarray(X, dtype=np.float32)
        N, T_, C = X.shape
        assert T_ == T and C <= Cm
        Xd = ctx.upload(X.reshape(N * T, C))
        ldd = _ceil(C, 4)
        hi, lo = ctx.empty((N * T, ldd)), ctx.empty((N * T, ldd))
        ctx.call('cpsd_split_tf32_2d', ptr(Xd), C, N * T, C, ptr(hi), ptr(lo), ldd)
        for u, t in enumerate((hi, lo)):
            _lib.check(ctx.lib.cpsd_tmap_encode_f32(ctypes.c_void_p(maps_h[2 * i + u].data_ptr()), ptr(t),
                                                    N * T, C, ldd, 128), 'tmap_encode')
        keep += [Xd, hi, lo]
    xmaps = maps_h.to(ctx.device, non_blocking=True)
    nprob = B * P
    nq = -(-Q // 32)
    ltc = 128 if Cm <= 128 else 256
    lthi, ltlo = ctx.zeros((nprob * nq, 32, ltc)), ctx.zeros((nprob * nq, 32, ltc))
    mul = ctx.zeros((nprob * nq, 32))
    mh = torch.zeros((2, 128), dtype=torch.uint8).pin_memory()
    for u, t in enumerate((lthi, ltlo)):
        _lib.check(ctx.lib.cpsd_tmap_encode_f32(ctypes.c_void_p(mh[u].data_ptr()), ptr(t), nprob * nq * 32,
                                                ltc, ltc, 32), 'tmap_encode')
    ltmaps = mh.to(ctx.device, non_blocking=True)
    Ld = ctx.upload(L)
    mud = ctx.upload(np.ascontiguousarray(mu, dtype=np.float32)) if mu is not None else None
    cdim = ctx.upload(np.tile(np.array([X.shape[2] for X in Xs], dtype=np.int32), B))
    ctx.call('cpsd_proj_tc_prep', ptr(Ld), Q, Cm * Q, ptr(mud), ptr(None), Cm, ptr(cdim), Q, ltc, ptr(lthi),
             ptr(ltlo), ptr(mul), nprob)
    dst_d = ctx.upload(np.ascontiguousarray(dst, dtype=np.int32))
    Z = ctx.zeros((B, n_rows, T * Q))
    ntr = np.array([X.shape[0] for X in Xs], dtype=np.int32)
    nch = np.array([X.shape[2] for X in Xs], dtype=np.int32)
    sms = torch.cuda.get_device_properties(ctx.device).multi_processor_count
    ctx.call('cpsd_proj_tc', ptr(xmaps), ptr(ltmaps), P, B, T, Q, ltc, ctypes.c_void_p(ntr.ctypes.data),
             ctypes.c_void_p(nch.ctypes.data), Nmax, ptr(dst_d), ptr(mul), ptr(Z), n_rows * T * Q, sms)
    out = Z.cpu().numpy()
    del keep
    return out


def class_mean(X, ids, device=None):
    """Mean over trials of each class.  X: (N, ...) ; ids: (N,) ints.  Classes in sorted id
    order.  Returns (classes, means (n_classes, ...))."""
    ctx = _ctx(device)
    X = np.asarray(X, dtype=np.float32)
    N = X.shape[0]
    TC = int(np.prod(X.shape[1:]))
    ids = np.asarray(ids)
    classes, inv = np.unique(ids, return_inverse=True)
    order = np.argsort(inv, kind='stable').astype(np.int32)
    mptr = np.concatenate([[0], np.cumsum(np.bincount(inv, minlength=len(classes)))])
    Xd = ctx.upload(np.ascontiguousarray(X.reshape(N, TC)))
    out = ctx.empty((len(classes), TC))
    pk = HostPack(ctx)
    o1, o2 = pk.add_ints(mptr), pk.add_ints(order)
    pk.reserve_ints()
    rec = np.zeros(1, dtype=_lib.CLASS_MEAN_DESC)
    rec[0] = (addr(Xd), pk.iaddr(o1), pk.iaddr(o2), addr(out), len(classes), TC, 0, 0)
    d = pk.add_descs(rec)
    pk.upload()
    ctx.call('cpsd_class_mean', pk.daddr(d), 1, len(classes), TC)
    return classes, out.cpu().numpy().reshape((len(classes),) + X.shape[1:])


def cca_solve(Saa, Sbb, Sab, device=None, check_rank=True):
    """CCA from scatter matrices (fp64 solve).  Returns dict(Ma, Mb, G (db x da, b->a), rho, info).
    Raises LinAlgError when a scatter matrix is numerically rank deficient (the reference
    truncates to ``matrix_rank`` there, AlignCCA.py:263-265; the Gram form cannot)."""
    ctx = _ctx(device)
    da, db = Saa.shape[0], Sbb.shape[0]
    dmax = _ceil(max(da, db), 4)
    S = np.zeros((2 * dmax, 2 * dmax), dtype=np.float64)
    S[:da, :da] = Saa
    S[dmax:dmax + db, dmax:dmax + db] = Sbb
    S[:da, dmax:dmax + db] = Sab
    Sd = ctx.upload(S)
    Ma, Mb, G = ctx.zeros((dmax, dmax)), ctx.zeros((dmax, dmax)), ctx.zeros((dmax, dmax))
    rho, info = ctx.zeros((dmax,)), ctx.zeros((4,), I32)
    rec = np.zeros(1, dtype=_lib.CCA_DESC)
    base = addr(Sd)
    rec[0] = (base, base + 8 * (dmax * 2 * dmax + dmax), base + 8 * dmax, 0, 0, addr(Ma), addr(Mb),
              addr(G), addr(rho), addr(info), da, db, 2 * dmax, dmax, dmax, 0, 1e-13, 0)
    pk = HostPack(ctx)
    pk.reserve_ints()
    d = pk.add_descs(rec)
    pk.upload()
    ws = ctx.empty((int(ctx.lib.cpsd_cca_solve_f64_ws_elems(1, dmax)),), F64)
    ctx.call('cpsd_cca_solve_f64', pk.daddr(d), 1, dmax, ptr(ws))
    inf = info.cpu().numpy()
    if check_rank and inf[1] == 1:
        raise np.linalg.LinAlgError('CCA: rank-deficient latent dynamics (fewer independent samples '
                                    'than latent dimensions, or duplicated dimensions)')
    dd = int(inf[0])
    return dict(Ma=Ma.cpu().numpy()[:da, :dd], Mb=Mb.cpu().numpy()[:db, :dd],
                G=G.cpu().numpy()[:db, :da], rho=rho.cpu().numpy()[:dd], info=inf)


def cca_solve_f32(Saa, Sbb, Sab, device=None):
    """The fp32 shared-memory solver (kept for comparison; d <= 116)."""
    ctx = _ctx(device)
    da, db = Saa.shape[0], Sbb.shape[0]
    dmax = _ceil(max(da, db), 4)
    S = np.zeros((2 * dmax, 2 * dmax), dtype=np.float32)
    S[:da, :da] = Saa
    S[dmax:dmax + db, dmax:dmax + db] = Sbb
    S[:da, dmax:dmax + db] = Sab
    Sd = ctx.upload(S)
    Ma, Mb, G = ctx.zeros((dmax, dmax)), ctx.zeros((dmax, dmax)), ctx.zeros((dmax, dmax))
    rho, info = ctx.zeros((dmax,)), ctx.zeros((4,), I32)
    rec = np.zeros(1, dtype=_lib.CCA_DESC)
    base = addr(Sd)
    rec[0] = (base, base + 4 * (dmax * 2 * dmax + dmax), base + 4 * dmax, 0, 0, addr(Ma), addr(Mb),
              addr(G), addr(rho), addr(info), da, db, 2 * dmax, dmax, dmax, 0, 1e-10, 0)
    pk = HostPack(ctx)
    pk.reserve_ints()
    d = pk.add_descs(rec)
    pk.upload()
    ctx.call('cpsd_cca_solve', pk.daddr(d), 1, dmax)
    inf = info.cpu().numpy()
    dd = int(inf[0])
    return dict(Ma=Ma.cpu().numpy()[:da, :dd], Mb=Mb.cpu().numpy()[:db, :dd],
                G=G.cpu().numpy()[:db, :da], rho=rho.cpu().numpy()[:dd], info=inf)


def svm_fit_ovr(S, y, C=1.0, dcd_epochs=0, max_newton=400, tol_newton=1e-9, tol_dcd=1e-4,
                device=None):
    """One-vs-rest L2-regularised squared-hinge linear SVM.  S: (n, k) features, y: (n,) ints.
    Returns (classes, W (n_classes, k+1) float64 with the bias last, info (n_classes, 4))."""
    ctx = _ctx(device)
    S = np.asarray(S, dtype=np.float32)
    n, k = S.shape
    y = np.asarray(y).astype(np.int32)
    classes = np.unique(y).astype(np.int32)
    lds = _ceil(n, 4)
    St = np.zeros((max(k, 1), lds), dtype=np.float32)
    St[:k, :n] = S.T
    Sd, yd = ctx.upload(St), ctx.upload(y)
    W = ctx.zeros((len(classes), k + 1), F64)
    info = ctx.zeros((len(classes), 4), I32)
    rec = np.zeros(len(classes), dtype=_lib.SVM_DESC)
    for c, cv in enumerate(classes):
        rec[c] = (addr(Sd), addr(yd), 0, addr(W, c * (k + 1)), addr(info, c * 4), n, k, lds,
                  int(cv), float(C), float(tol_dcd), float(tol_newton), int(max_newton),
                  int(dcd_epochs))
    pk = HostPack(ctx)
    pk.reserve_ints()
    d = pk.add_descs(rec)
    pk.upload()
    ctx.call('cpsd_svm_fit_ovr_ex', pk.daddr(d), len(classes), k, lds, int(dcd_epochs))
    return classes, W.cpu().numpy(), info.cpu().numpy()


def svm_predict_ovr(S, classes, W, device=None, return_decision=False):
    ctx = _ctx(device)
    S = np.asarray(S, dtype=np.float32)
    n, k = S.shape
    Xt = ctx.upload(np.ascontiguousarray(S.T) if k else np.zeros((1, n), dtype=np.float32))
    Wd = ctx.upload(np.ascontiguousarray(W, dtype=np.float64))
    cd = ctx.upload(np.asarray(classes, dtype=np.int32))
    yh = ctx.empty((n,), I32)
    dec = ctx.empty((n, len(classes)), F64)
    ctx.call('cpsd_svm_predict_ovr', ptr(Xt), n, 0, ptr(Wd), k + 1, 0, ptr(None), k, ptr(None), n,
             ptr(cd), len(classes), ptr(yh), ptr(dec), 1)
    if return_decision:
        return yh.cpu().numpy(), dec.cpu().numpy()
    return yh.cpu().numpy()


_SVC_KERNELS = {'linear': 0, 'rbf': 1}


def svc_fit_ovo(S, y, C=1.0, kernel='rbf', gamma='scale', balanced=False, tol=1e-3,
                max_iter=1000000, device=None):
    """libsvm-style C-SVC, one-vs-one (sklearn.svm.SVC as the reference scripts build it,
    scripts/aligned_decode_svm_ncv.py:313-317).  S: (n, k) features, y: (n,) ints.  Returns a
    dict holding the device state ``predict`` needs plus host copies of the fitted quantities
    (coef in libsvm's sv_coef layout (n_classes-1, n), rho per pair, gamma, info)."""
    ctx = _ctx(device)
    S = np.asarray(S, dtype=np.float32)
    n, k = S.shape
    y = np.asarray(y).astype(np.int32)
    classes, counts = np.unique(y, return_counts=True)
    classes = classes.astype(np.int32)
    ncls = len(classes)
    if ncls < 2:
        raise ValueError('The number of classes has to be greater than one; got %d class' % ncls)
    if kernel not in _SVC_KERNELS:
        raise ValueError("kernel must be 'linear' or 'rbf'")
    if isinstance(gamma, str):
        if gamma == 'scale':
            g = -1.0
        elif gamma == 'auto':
            g = 1.0 / max(k, 1)
        else:
            raise ValueError("gamma must be 'scale', 'auto' or a positive float")
    else:
        g = float(gamma)
        if g <= 0:
            raise ValueError('gamma must be positive')
    lds = _ceil(n, 4)
    St = np.zeros((max(k, 1), lds), dtype=np.float32)
    St[:k, :n] = S.T
    Sd, yd, cd = ctx.upload(St), ctx.upload(y), ctx.upload(classes)
    K = ctx.empty((n, lds))
    gam = ctx.empty((1,), F64)
    perm, off, sqn = ctx.empty((lds,), I32), ctx.empty((ncls + 1,), I32), ctx.empty((lds,), F64)
    kid = _SVC_KERNELS[kernel]
    ctx.call('cpsd_svc_kernel_matrix', ptr(Sd), lds, 0, ptr(None), k, ptr(None), n, n, ptr(yd), 0, ptr(cd),
             ncls, kid, g, ptr(gam), ptr(perm), ptr(off), ptr(sqn), ptr(K), lds, 0, 1)
    npair = ncls * (ncls - 1) // 2
    coef = ctx.empty((ncls - 1, lds), F64)
    rho = ctx.empty((npair,), F64)
    info = ctx.empty((npair, 2), I32)
    srt = np.sort(counts)
    m_max = int(srt[-1] + srt[-2])
    ctx.call('cpsd_svc_fit_ovo', ptr(K), lds, 0, ptr(perm), ptr(off), ptr(None), n, ncls, float(C),
             int(bool(balanced)), float(tol), int(max_iter), ptr(coef), lds, ptr(rho), ptr(info), m_max, 1)
    return dict(St=Sd, y=yd, classes_dev=cd, coef_dev=coef, rho_dev=rho, gamma_dev=gam, n=n, k=k, lds=lds,
                kernel=kid, classes=classes, coef=coef.cpu().numpy()[:, :n], rho=rho.cpu().numpy(),
                gamma=float(gam.cpu().numpy()[0]), info=info.cpu().numpy())


def svc_predict_ovo(model, S, device=None, return_decision=False):
    ctx = _ctx(device)
    S = np.asarray(S, dtype=np.float32)
    nt, k = S.shape
    assert k == model['k'], 'feature count differs from fit'
    ncls = len(model['classes'])
    npair = ncls * (ncls - 1) // 2
    Zt = ctx.upload(np.ascontiguousarray(S.T) if k else np.zeros((1, nt), dtype=np.float32))
    yh = ctx.empty((nt,), I32)
    dec = ctx.empty((nt, npair), F64)
    n, lds = model['n'], model['lds']
    ctx.call('cpsd_svc_predict_ovo', ptr(model['St']), lds, 0, ptr(Zt), nt, 0, ptr(None), k, ptr(None), n, n,
             ptr(None), nt, ptr(model['y']), 0, ptr(model['classes_dev']), ncls, model['kernel'],
             ptr(model['gamma_dev']), ptr(model['coef_dev']), lds, ptr(model['rho_dev']), ptr(yh),
             ptr(dec), k, 1)
    if return_decision:
        return yh.cpu().numpy(), dec.cpu().numpy()
    return yh.cpu().numpy()
