"""Kernel-level parity: every CUDA kernel (through the C-ABI, host buffers in/out) against
a float64 numpy statement of the same operation on seeded inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(lib_built):
    from cross_patient_speech_decoding_b200 import ops
    return ops


def _rand_sym(rng, n, psd=True):
    X = rng.standard_normal((n, max(n // 2, 3) if psd else n))
    return X @ X.T if psd else X + X.T


@pytest.mark.parametrize('n', [1, 2, 3, 13, 30, 69, 100, 127, 128])
def test_eig_small(ops, n):
    rng = np.random.default_rng(n)
    A = np.stack([_rand_sym(rng, n, psd=(i % 2 == 0)) for i in range(5)])
    ev, V = ops.eig_sym(A)
    for i in range(5):
        ref = np.linalg.eigvalsh(A[i])[::-1]
        scale = np.abs(ref).max() + 1e-30
        assert np.abs(ev[i] - ref).max() <= 2e-5 * scale
        assert np.abs(V[i].T @ V[i] - np.eye(n)).max() < 1e-4
        assert np.abs(A[i] @ V[i] - V[i] * ev[i]).max() <= 5e-5 * scale


def test_eig_small_ragged_sizes(ops):
    rng = np.random.default_rng(0)
    ns = np.array([5, 40, 17, 64], dtype=np.int32)
    A = np.zeros((4, 64, 64))
    for i, n in enumerate(ns):
        A[i, :n, :n] = _rand_sym(rng, n)
    ev, V = ops.eig_sym(A, n=ns)
    for i, n in enumerate(ns):
        ref = np.linalg.eigvalsh(A[i, :n, :n])[::-1]
        assert np.abs(ev[i, :n] - ref).max() <= 2e-5 * ref[0]


@pytest.mark.parametrize('tc', [False, True])
@pytest.mark.parametrize('n', [129, 240, 300, 640])
def test_eig_block(ops, n, tc):
    rng = np.random.default_rng(n)
    A = np.stack([_rand_sym(rng, n, psd=(i == 0)) for i in range(2)])
    ev, V, sw = ops.eig_sym(A, return_sweeps=True, tensor_cores=tc)
    assert (sw <= 15).all(), sw
    for i in range(2):
        ref = np.linalg.eigvalsh(A[i])[::-1]
        scale = np.abs(ref).max()
        assert np.abs(ev[i] - ref).max() <= 6e-5 * scale
        assert np.abs(V[i].T @ V[i] - np.eye(n)).max() < 1e-4, np.abs(V[i].T @ V[i] - np.eye(n)).max()
        assert np.abs(A[i] @ V[i] - V[i] * ev[i]).max() <= 1e-4 * scale


def test_eig_block_pooled_spectrum(ops):
    """A spectrum like the pooled Gram: a few large, a flat bulk, exact rank deficiency."""
    rng = np.random.default_rng(1)
    n, F = 700, 900
    Z = rng.standard_normal((n, F)) + 4 * rng.standard_normal((n, 6)) @ rng.standard_normal((6, F))
    Z -= Z.mean(0)
    K = Z @ Z.T
    ev, V = ops.eig_sym(K, tensor_cores=True)
    ref_w, ref_v = np.linalg.eigh(K)
    ref_w, ref_v = ref_w[::-1], ref_v[:, ::-1]
    assert np.abs(ev - ref_w).max() <= 3e-5 * ref_w[0]
    k = 6                              # the well-separated signal subspace
    s = np.linalg.svd(ref_v[:, :k].T @ V[:, :k], compute_uv=False)
    assert s.min() > 1 - 5e-6
    assert np.abs(K @ V - V * ev).max() <= 1e-4 * ref_w[0]   # residual of every eigen-pair
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-4


def test_select_k_modes(ops):
    ev = np.array([[5.0, 3.0, 1.0, 0.5, 0.5, 0.0]], dtype=np.float32)
    r = np.cumsum(ev[0]) / ev[0].sum()
    for thr in (0.45, 0.85, 0.93):
        assert ops.select_k(ev, thr, 0)[0] == np.searchsorted(r, thr, side='right') + 1
        assert ops.select_k(ev, thr, 1)[0] == np.argmax(r > thr)
        assert ops.select_k(ev, thr, 2)[0] == np.argmax(r >= thr) + 1
    assert ops.select_k(ev, 0.999999, 1)[0] == np.argmax(r > 0.999999)
    assert ops.select_k(ev, 4, 3)[0] == 4


@pytest.mark.parametrize('shape', [(1000, 128, 128), (333, 33, 70), (64, 201, 201), (5, 7, 3)])
def test_gram_tn(ops, shape):
    rows, p, q = shape
    rng = np.random.default_rng(rows)
    A = rng.standard_normal((rows, p)) + 0.5
    B = rng.standard_normal((rows, q)) - 0.25
    G = ops.gram_tn(A, B)
    ref = A.T @ B
    assert np.abs(G - ref).max() <= 2e-5 * np.abs(ref).max() + 1e-4
    muA, muB = A.mean(0), B.mean(0)
    Gc = ops.gram_tn(A, B, muA=muA, muB=muB, alpha=0.5)
    refc = 0.5 * (A - muA).T @ (B - muB)
    assert np.abs(Gc - refc).max() <= 2e-5 * np.abs(refc).max() + 1e-4
    if p == q:
        Gs = ops.gram_tn(A)
        assert np.abs(Gs - A.T @ A).max() <= 2e-5 * np.abs(A.T @ A).max()
        assert np.array_equal(Gs, Gs.T)


def test_gram_tn_segments(ops):
    rng = np.random.default_rng(3)
    T, C = 50, 40
    X = rng.standard_normal((12 * T, C))
    sel = np.array([7, 0, 3, 11])
    G = ops.gram_tn(X, seg_rows=sel * T, seg_len=T)
    Xs = np.concatenate([X[s * T:(s + 1) * T] for s in sel])
    assert np.abs(G - Xs.T @ Xs).max() <= 2e-5 * np.abs(Xs.T @ Xs).max()
    m = ops.colmean(X, seg_rows=sel * T, seg_len=T)
    assert np.abs(m - Xs.mean(0)).max() < 1e-5


@pytest.mark.parametrize('shape', [(259, 259, 2600), (130, 130, 515), (8, 300, 1000), (1, 1, 5)])
def test_gram_nt(ops, shape):
    m, n, k = shape
    rng = np.random.default_rng(m + k)
    A = rng.standard_normal((m, k))
    if m == n:
        K = ops.gram_nt(A)
        ref = A @ A.T
        assert np.array_equal(K, K.T)
    else:
        B = rng.standard_normal((n, k))
        K = ops.gram_nt(A, B)
        ref = A @ B.T
    assert np.abs(K - ref).max() <= 3e-5 * np.abs(ref).max() + 1e-5


@pytest.mark.parametrize('C,q', [(128, 30), (33, 13), (201, 100), (64, 128), (10, 1)])
def test_project(ops, C, q):
    rng = np.random.default_rng(C + q)
    X = rng.standard_normal((7, 53, C))
    W = rng.standard_normal((C, q))
    mu = rng.standard_normal(C)
    Y = ops.project(X, W, mu)
    ref = (X - mu) @ W
    assert Y.shape == ref.shape
    assert np.abs(Y - ref).max() <= 2e-5 * np.abs(ref).max()
    Y0 = ops.project(X, W)
    assert np.abs(Y0 - X @ W).max() <= 2e-5 * np.abs(X @ W).max()


def test_class_mean(ops):
    rng = np.random.default_rng(5)
    X = rng.standard_normal((40, 20, 17))        # T*C = 340 (float4 path), odd C below
    ids = rng.integers(0, 9, 40)
    cls, M = ops.class_mean(X, ids)
    assert np.array_equal(cls, np.unique(ids))
    for i, c in enumerate(cls):
        assert np.abs(M[i] - X[ids == c].mean(0)).max() < 1e-5
    X2 = rng.standard_normal((15, 7, 3))         # T*C = 21: scalar path
    ids2 = rng.integers(0, 4, 15)
    cls2, M2 = ops.class_mean(X2, ids2)
    for i, c in enumerate(cls2):
        assert np.abs(M2[i] - X2[ids2 == c].mean(0)).max() < 1e-5


def _cca_ref(La, Lb):
    """Reference CCA_align in float64 (AlignCCA.py:235-285), written out for the test."""
    La = La - La.mean(0)
    Lb = Lb - Lb.mean(0)
    Qa, Ra = np.linalg.qr(La)
    Qb, Rb = np.linalg.qr(Lb)
    U, S, Vt = np.linalg.svd(Qa.T @ Qb)
    d = min(La.shape[1], Lb.shape[1])
    Ma = np.linalg.pinv(Ra) @ U[:, :d]
    Mb = np.linalg.pinv(Rb) @ Vt.T[:, :d]
    return Ma, Mb, np.clip(S[:d], 0, 1), La, Lb


@pytest.mark.parametrize('da,db', [(13, 13), (20, 14), (9, 26), (60, 60), (100, 97), (1, 1), (116, 116),
                                   (130, 101), (88, 150), (201, 180)])
def test_cca_solve(ops, da, db):
    rng = np.random.default_rng(da * 100 + db)
    n = 900
    Zs = rng.standard_normal((n, max(da, db)))
    La = Zs[:, :da] @ rng.standard_normal((da, da)) + 0.3 * rng.standard_normal((n, da))
    Lb = Zs[:, :db] @ rng.standard_normal((db, db)) + 0.3 * rng.standard_normal((n, db))
    Ma, Mb, rho, Lac, Lbc = _cca_ref(La, Lb)
    out = ops.cca_solve(Lac.T @ Lac, Lbc.T @ Lbc, Lac.T @ Lbc)
    assert out['info'][0] == min(da, db)
    assert out['info'][1] == 0
    assert np.abs(out['rho'] - rho).max() < 2e-6        # fp64 solve, fp32 output
    # canonical variates: correlations of the projected data equal rho, variates are white
    Pa, Pb = Lac @ out['Ma'], Lbc @ out['Mb']
    assert np.abs(Pa.T @ Pa - np.eye(len(rho))).max() < 1e-4
    assert np.abs(np.diag(Pa.T @ Pb) - rho).max() < 1e-4
    # b -> a map equals M_b pinv(M_a) of the reference
    Gref = Mb @ np.linalg.pinv(Ma)
    assert np.abs(out['G'] - Gref).max() <= 2e-5 * np.abs(Gref).max()


@pytest.mark.parametrize('da,db', [(30, 30), (100, 97)])
def test_cca_solve_fp64_beats_fp32_on_ill_conditioned_latents(ops, da, db):
    """Latents whose variances span 1e5 (PCA scores down to the noise floor): the fp64 solver keeps
    the b->a map at fp32-output accuracy where the fp32 shared-memory solver loses digits."""
    rng = np.random.default_rng(da)
    n = 2000
    sc = np.logspace(0, -2.5, max(da, db))
    Zs = rng.standard_normal((n, max(da, db)))
    La = (Zs[:, :da] + 0.5 * rng.standard_normal((n, da))) * sc[:da]
    Lb = (Zs[:, :db] @ np.linalg.qr(rng.standard_normal((db, db)))[0] + 0.5 * rng.standard_normal((n, db))) * sc[:db]
    Ma, Mb, rho, Lac, Lbc = _cca_ref(La, Lb)
    Gref = Mb @ np.linalg.pinv(Ma)
    S = (Lac.T @ Lac, Lbc.T @ Lbc, Lac.T @ Lbc)
    e64 = np.abs(Lbc @ ops.cca_solve(*S)['G'] - Lbc @ Gref).max() / np.abs(Lbc @ Gref).max()
    e32 = np.abs(Lbc @ ops.cca_solve_f32(*S)['G'] - Lbc @ Gref).max() / np.abs(Lbc @ Gref).max()
    assert e64 < 1e-5, e64
    assert e64 < e32


@pytest.mark.parametrize('n', [129, 150, 201, 256])
def test_eig_sym_f64_above_128(ops, n):
    """fp64 one-sided Jacobi eigen-solver for 128 < n <= 256 (patients with more than 128
    channels) against numpy.linalg.eigh: spectrum, orthonormality, residual, sorted order."""
    rng = np.random.default_rng(n)
    X = rng.standard_normal((3 * n, n)) * np.logspace(0, -2, n)
    A = np.stack([X.T @ X, _rand_sym(rng, n)])
    ev, V, sw = ops.eig_sym(A, f64=True, return_sweeps=True)
    assert (sw < 40).all(), sw
    for i in range(2):
        ref = np.linalg.eigvalsh(A[i])[::-1]
        scale = np.abs(ref).max()
        assert np.abs(ev[i] - ref).max() <= 2e-7 * scale
        assert np.abs(V[i].T @ V[i] - np.eye(n)).max() < 2e-6
        assert np.abs(A[i] @ V[i] - V[i] * ev[i]).max() <= 2e-6 * scale


def test_svm_matches_liblinear_primal(ops):
    from sklearn.svm import LinearSVC
    rng = np.random.default_rng(0)
    n, k = 300, 20
    X = rng.standard_normal((n, k)) * np.linspace(60, 5, k)     # unscaled PCA-like scores
    w = rng.standard_normal((k, 4))
    y = np.argmax(X @ w + 30 * rng.standard_normal((n, 4)), 1) + 3
    cls, W, info = ops.svm_fit_ovr(X, y, C=1.0, dcd_epochs=2)
    assert (info[:, 3] == 0).all(), info
    ref = LinearSVC(dual=False, C=1.0, tol=1e-12, max_iter=100000).fit(X, y)
    Wref = np.hstack([ref.coef_, ref.intercept_[:, None]])
    assert np.array_equal(cls, ref.classes_)
    assert np.abs(W - Wref).max() <= 1e-6 * np.abs(Wref).max() + 1e-9
    Xte = rng.standard_normal((50, k)) * np.linspace(60, 5, k)
    yh, dec = ops.svm_predict_ovr(Xte, cls, W, return_decision=True)
    assert np.array_equal(yh, ref.predict(Xte))
    assert np.abs(dec - ref.decision_function(Xte)).max() < 1e-4


def test_svm_dcd_alone_matches_liblinear_dual(ops):
    """Phase 1 alone (no Newton) on a well-conditioned problem where liblinear's own dual CD
    converges: same optimum as LinearSVC(dual=True)."""
    from sklearn.svm import LinearSVC
    rng = np.random.default_rng(1)
    n, k = 200, 10
    X = rng.standard_normal((n, k))
    y = (X[:, 0] + 0.5 * X[:, 1] + 0.7 * rng.standard_normal(n) > 0).astype(int)
    cls, W, info = ops.svm_fit_ovr(X, y, C=0.5, dcd_epochs=5000, max_newton=0, tol_dcd=1e-7)
    assert (info[:, 3] == 0).all(), info
    ref = LinearSVC(dual=True, C=0.5, tol=1e-8, max_iter=200000, random_state=0).fit(X, y)
    wref = np.hstack([ref.coef_[0], ref.intercept_])
    # binary sklearn keeps one row for the positive class (classes_[1])
    assert np.abs(W[1] - wref).max() < 1e-4
    assert np.abs(W[0] + wref).max() < 1e-4


def test_svm_edge_cases(ops):
    rng = np.random.default_rng(2)
    X = rng.standard_normal((30, 1))
    y = np.array([1] * 15 + [2] * 15)
    cls, W, info = ops.svm_fit_ovr(X, y)
    assert W.shape == (2, 2) and np.isfinite(W).all()
    # k = 0 features: bias only
    cls, W, info = ops.svm_fit_ovr(np.zeros((10, 0)), np.array([0, 1] * 5))
    assert W.shape == (2, 1) and np.isfinite(W).all()


@pytest.mark.parametrize('shape', [(128, 64), (259, 2600), (700, 1000), (1152, 6000)])
def test_gram_nt_tensor_core(ops, shape):
    """tcgen05 3xTF32 Gram against float64, and against the fp32 SIMT kernel."""
    m, k = shape
    rng = np.random.default_rng(m + k)
    A = rng.standard_normal((m, k)) * (1 + 5 * rng.random((m, 1)))
    K = ops.gram_nt(A, tensor_cores=True)
    ref = A @ A.T
    err = np.abs(K - ref).max() / np.abs(ref).max()
    assert err <= 5e-6, err
    assert np.array_equal(K, K.T)
    Ks = ops.gram_nt(A)
    assert np.abs(K - Ks).max() / np.abs(ref).max() <= 1e-5


def test_eig_f64_resolves_noise_level_directions(ops):
    """PCA-like spectrum: a few large eigenvalues over a nearly flat noise floor.  The fp64
    matrix iteration must recover individual noise-level eigenvectors (gap-sensitive), which
    is what keeps canonical correlations of noise dimensions within 1e-4 of the reference."""
    rng = np.random.default_rng(7)
    n = 128
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([np.linspace(8, 1, 12), 0.0225 * (1 + 0.3 * rng.random(n - 12))])
    A = (Q * lam) @ Q.T
    ev, V = ops.eig_sym(A, f64=True)
    w, U = np.linalg.eigh(A)
    w, U = w[::-1], U[:, ::-1]
    assert np.abs(ev - w).max() <= 1e-6 * w[0]
    cos = np.abs(np.sum(U * V, axis=0))           # per-eigenvector alignment
    assert cos.min() > 1 - 5e-6, cos.min()
    assert np.abs(V.T @ V - np.eye(n)).max() < 2e-5


@pytest.mark.parametrize('trans_a', [False, True])
@pytest.mark.parametrize('shape', [(128, 128, 128), (1, 128, 128), (70, 33, 129), (5, 3, 2)])
def test_dgemm_batched(ops, shape, trans_a):
    m, n, k = shape
    rng = np.random.default_rng(m + n + k)
    A = rng.standard_normal((4, k, m) if trans_a else (4, m, k))
    B = rng.standard_normal((3, k, n))
    D = rng.standard_normal((6, m, n))
    ai, bi = [3, 0, 0, 2, 1, 3], [2, 2, 0, 1, 0, 1]
    C = ops.dgemm_batched(A, B, trans_a=trans_a, alpha=-0.5, beta=1.5, D=D, a_idx=ai, b_idx=bi)
    for p in range(6):
        a = A[ai[p]].T if trans_a else A[ai[p]]
        ref = -0.5 * a @ B[bi[p]] + 1.5 * D[p]
        assert np.abs(C[p] - ref).max() <= 1e-13 * (np.abs(ref).max() + 1)


@pytest.mark.parametrize('n', [128, 111, 37])
def test_eig_warm_start_same_pairs_fewer_sweeps(ops, n):
    """Fold-like perturbations of one scatter matrix, solved cold and warm-started from the
    eigenvectors of the unperturbed matrix: same eigenpairs (to the fp32 accumulator's accuracy),
    clearly fewer sweeps; a problem flagged cold inside the warm launch matches the cold solver;
    sel / out_idx route a subset of the batch to chosen output slots."""
    rng = np.random.default_rng(n)
    X = rng.standard_normal((140, 40, 12)) @ rng.standard_normal((12, n)) + 0.3 * rng.standard_normal((140, 40, n))
    G = np.einsum('ntc,ntd->ncd', X, X)
    A = np.stack([G.sum(0) - G[rng.permutation(140)[:7]].sum(0) for _ in range(6)])
    _, V0 = ops.eig_sym(G.sum(0), f64=True)
    ev_c, V_c, sw_c = ops.eig_sym(A, f64=True, return_sweeps=True)
    vi = np.array([0, 0, 0, -1, 0, 0])
    ev_w, V_w, sw_w = ops.eig_sym_warm(A, V0[None], v0_idx=vi)
    for i in range(6):
        w, U = np.linalg.eigh(A[i])
        w, U = w[::-1], U[:, ::-1]
        assert np.abs(ev_w[i] - w).max() <= 1e-6 * w[0]
        assert np.abs(V_w[i].T @ V_w[i] - np.eye(n)).max() < 2e-5
        assert np.abs(A[i] @ V_w[i] - V_w[i] * w).max() <= 2e-5 * w[0]
        gaps = np.minimum(np.abs(np.diff(w, prepend=np.inf)), np.abs(np.diff(w, append=-np.inf)))
        ok = gaps > 1e-3 * w[0]                     # well-separated pairs: same vectors up to sign
        assert np.abs(np.abs(np.sum(U * V_w[i], axis=0)) - 1)[ok].max() < 1e-5
    assert np.abs(ev_w[3] - ev_c[3]).max() <= 1e-6 * ev_c[3, 0]
    assert sw_w[vi >= 0].max() < sw_c[:6].min(), (sw_w, sw_c)
    # subset + output slots
    ev_s, V_s, _ = ops.eig_sym_warm(A, V0[None], v0_idx=np.zeros(6), sel=[4, 1], out_idx=[9, 2, 9, 9, 0, 9],
                                    n_out=3)
    assert np.abs(ev_s[0] - ev_w[4]).max() <= 1e-6 * ev_w[4, 0]
    assert np.abs(ev_s[2] - ev_w[1]).max() <= 1e-6 * ev_w[1, 0]
    assert np.abs(ev_s[1]).max() == 0


def test_gram_tn_f64_accumulation(ops):
    import torch
    from cross_patient_speech_decoding_b200 import _lib
    from cross_patient_speech_decoding_b200.device import Context, HostPack, addr
    ctx = Context.get(None)
    rng = np.random.default_rng(9)
    A = (rng.standard_normal((20000, 96)) + 3.0).astype(np.float32)
    Ad = ctx.upload(A)
    out = ctx.zeros((96, 96), torch.float64)
    pk = HostPack(ctx)
    o = pk.add_ints([0])
    pk.reserve_ints()
    rec = np.zeros(1, dtype=_lib.GRAM_TN_DESC)
    rec[0] = (addr(Ad), addr(Ad), pk.iaddr(o), pk.iaddr(o), 0, 0, addr(out), 1, 20000, 96, 96, 96,
              96, 96, 1, 1.0, 0)
    d = pk.add_descs(rec)
    pk.upload()
    ctx.call('cpsd_gram_tn_f64', pk.daddr(d), 1, 96, 96)
    G = out.cpu().numpy()
    ref = A.astype(np.float64).T @ A.astype(np.float64)
    assert np.abs(G - ref).max() <= 1e-12 * np.abs(ref).max()


# ------------------------------------------------------------------ subspace (top-k) solver
@pytest.mark.parametrize('trans_a', [False, True])
@pytest.mark.parametrize('shape', [(128, 64, 16), (1152, 128, 1152), (130, 70, 37), (5, 3, 2),
                                   (256, 128, 128)])
def test_sgemm_batched(ops, shape, trans_a):
    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((3, K, M) if trans_a else (3, M, K))
    B = rng.standard_normal((3, K, N))
    C = ops.sgemm_batched(A, B, trans_a=trans_a, alpha=0.5)
    ref = 0.5 * np.einsum('bkm,bkn->bmn' if trans_a else 'bmk,bkn->bmn', A, B)
    assert np.abs(C - ref).max() <= 2e-6 * np.sqrt(K) * np.abs(ref).max() + 1e-6


@pytest.mark.parametrize('m', [4, 33, 128])
def test_chol_inv(ops, m):
    rng = np.random.default_rng(m)
    Y = rng.standard_normal((4, 3 * m + 5, m)) * np.linspace(1.0, 30.0, m)
    S = np.einsum('bkm,bkn->bmn', Y, Y)
    Rinv, st = ops.chol_inv(S)
    assert not st.any()
    for b in range(4):
        R = np.linalg.cholesky(S[b]).T
        ref = np.linalg.inv(R)
        assert np.abs(Rinv[b] - ref).max() <= 1e-5 * np.abs(ref).max()
        assert np.abs(np.tril(Rinv[b], -1)).max() == 0.0
    # rank-deficient Gram is flagged, output stays finite
    S[1] = np.outer(np.ones(m), np.ones(m))
    Rinv, st = ops.chol_inv(S)
    assert st[1] == 1 and st[0] == 0 and np.isfinite(Rinv).all()


def _pooled_like(rng, n, F, decay):
    X = rng.standard_normal((n, F)) * decay[None, :F]
    X -= X.mean(axis=0)
    return X @ X.T


@pytest.mark.parametrize('tc', [False, True])
@pytest.mark.parametrize('n', [300, 1144])
def test_eig_topk_leading_pairs(ops, n, tc):
    rng = np.random.default_rng(n)
    decay = 1.0 / (1.0 + np.arange(4000) / 20.0)
    A = np.stack([_pooled_like(rng, n, 4000, decay) for _ in range(3)])
    out = ops.eig_topk(A, m=128, iters=8, tensor_cores=tc)
    assert not out['status'].any()
    for b in range(3):
        w, U = np.linalg.eigh(A[b])
        w, U = w[::-1], U[:, ::-1]
        assert abs(out['total'][b] - np.trace(A[b])) <= 1e-5 * np.trace(A[b])
        k = 64
        assert np.abs(out['evals'][b, :k] - w[:k]).max() <= 2e-5 * w[0]
        V = out['V'][b][:, :k]
        assert np.abs(V.T @ V - np.eye(k)).max() < 1e-4
        assert out['resid'][b, :k].max() <= 2e-5 * w[0]
        # same invariant subspace as LAPACK: principal angles ~ 0
        sv = np.linalg.svd(U[:, :k].T @ V, compute_uv=False)
        assert sv.min() > 1 - 1e-6


def test_eig_topk_resume_and_ragged(ops):
    rng = np.random.default_rng(5)
    ns = np.array([260, 300], dtype=np.int32)
    decay = 1.0 / (1.0 + np.arange(2000) / 60.0)      # slow decay: one round is not enough
    A = np.zeros((2, 300, 300))
    for b, n in enumerate(ns):
        A[b, :n, :n] = _pooled_like(rng, n, 2000, decay)
        A[b, n:, :] = 7.0                              # garbage in the padding must be ignored
        A[b, :, n:] = 7.0
    one = ops.eig_topk(A, m=64, iters=2, rounds=1, n=ns)
    more = ops.eig_topk(A, m=64, iters=2, rounds=6, n=ns)
    for b, n in enumerate(ns):
        w = np.linalg.eigvalsh(A[b, :n, :n])[::-1]
        assert more['resid'][b, :20].max() < one['resid'][b, :20].max()
        assert np.abs(more['evals'][b, :20] - w[:20]).max() <= 2e-5 * w[0]
        if n < A.shape[1]:
            assert np.abs(more['V'][b][n:, :]).max() == 0.0


def test_select_k_total(ops):
    ev = np.array([[5.0, 3.0, 1.0, 0.5]], dtype=np.float32)
    # mode 0: k = #{cumulative ratio <= thr} + 1
    assert ops.select_k(ev, 0.82, 0, total=[10.0])[0] == 3    # ratios .5 .8 .9 .95 of the given total
    assert ops.select_k(ev, 0.82, 0)[0] == 2                  # ratios .526 .842 ... of their own sum


@pytest.mark.parametrize('shape', [
    dict(n=(20, 31, 9), c=(128, 96, 64), t=25, q=30, b=3),
    dict(n=(40, 17), c=(128, 128), t=200, q=32, b=5),
    dict(n=(12, 14, 5, 8), c=(36, 4, 100, 128), t=7, q=7, b=2),
    dict(n=(11, 9, 14), c=(63, 111, 74), t=13, q=30, b=3),          # odd channel counts (padded row stride)
    dict(n=(10, 12, 7), c=(149, 201, 111), t=19, q=30, b=3),        # more than 128 channels: two panels
    dict(n=(9, 13), c=(256, 171), t=16, q=12, b=2),
    dict(n=(14, 10), c=(128, 96), t=11, q=60, b=3),                 # wide latents: chunks of 32
    dict(n=(8, 12, 6), c=(201, 144, 63), t=9, q=100, b=2),          # both
    dict(n=(7, 9), c=(200, 128), t=5, q=128, b=2),
])
def test_proj_tc_direct(ops, shape):
    """k_proj_tc (tcgen05 3xTF32 pooled projection) on its own against fp64 (X - mu) L: ragged
    trial counts, channel counts below 128, odd time axis, destination tables with skipped
    trials and per-fold permutations."""
    rng = np.random.default_rng(3)
    P, B, T, Q = len(shape['n']), shape['b'], shape['t'], shape['q']
    Cm = max(shape['c'])
    Xs = [rng.standard_normal((n, T, c)) * rng.uniform(0.5, 3.0, c) + rng.standard_normal(c)
          for n, c in zip(shape['n'], shape['c'])]
    L = rng.standard_normal((B, P, Cm, Q))
    mu = rng.standard_normal((B, P, Cm))
    for v, c in enumerate(shape['c']):
        L[:, v, c:, :] = 0.0
    Nmax = max(shape['n'])
    ntot = sum(shape['n'])
    dst = -np.ones((B, P, Nmax), dtype=np.int32)
    for f in range(B):
        perm = rng.permutation(ntot)
        o = 0
        for v, n in enumerate(shape['n']):
            dst[f, v, :n] = perm[o:o + n]
            o += n
        dst[f, 0, f % shape['n'][0]] = -1           # one skipped trial per fold
    Z = ops.project_pool_tc(Xs, L, mu, dst, ntot)
    Z = Z.reshape(B, ntot, T, Q)
    worst = 0.0
    for f in range(B):
        used = np.zeros(ntot, dtype=bool)
        for v, n in enumerate(shape['n']):
            c = shape['c'][v]
            ref = (Xs[v].astype(np.float32).astype(np.float64) -
                   mu[f, v, :c].astype(np.float32).astype(np.float64)) @ \
                L[f, v, :c].astype(np.float32).astype(np.float64)
            for tr in range(n):
                d = dst[f, v, tr]
                if d < 0:
                    continue
                used[d] = True
                err = np.abs(Z[f, d] - ref[tr]).max() / np.abs(ref).max()
                worst = max(worst, err)
        assert not Z[f, ~used].any()                 # rows nobody maps to stay untouched
    assert worst < 2e-5, worst      # 3xTF32 (one TF32 pass would be ~1e-3)
