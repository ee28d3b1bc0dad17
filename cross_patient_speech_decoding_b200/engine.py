"""Batched cross-validated align -> reduce -> decode engine.

One call processes a whole list of CV folds ("units") for one target patient and a set
of cross patients: every fold fits the alignment (MCCA / pairwise CCA / none), projects
all trials into the shared latent space, pools them, runs the decoder-stage PCA on the
pooled trials x (time*latent) matrix and trains + scores the one-vs-rest linear SVM.
This is the loop body of the reference's scripts/aligned_decode_svm_ncv.py:344-442 with
``crossPtDecoder_{mcca,sepAlign,sepDimRed}`` (decoders/cross_pt_decoders.py) and the
``make_pipeline(DimRedReshape(PCA), <linear SVM>)`` decoder, executed for all folds at
once by the kernels of libcpsd_b200.so.  Patient data stays resident in HBM; per batch
the host only uploads fold index tables + descriptor records and downloads predictions.
"""
import ctypes
import os
import time

import numpy as np
import torch

from . import _lib
from .device import Context, HostPack, LaneContext, addr, ptr
from .folds import class_ids

F32 = torch.float32
I32 = torch.int32


def _ceil(a, b):
    return (a + b - 1) // b * b


_LANE_STREAMS = {}
# idle wait of the lane schedulers between polls of their CUDA events (seconds; 0 = yield only)
_IDLE_SLEEP = float(os.environ.get('CPSD_IDLE_SLEEP', '0'))


def _lane_stream(device, i):
    """One CUDA stream per (device, lane index) for the whole process: torch's caching allocator
    pools memory per stream, so engines created one after the other reuse the same blocks."""
    key = (str(device), i)
    if key not in _LANE_STREAMS:
        _LANE_STREAMS[key] = torch.cuda.Stream(device)
    return _LANE_STREAMS[key]


class View:
    """One patient resident on the device."""

    def __init__(self, ctx, X, y, y_align, cls_ids, upload_stream=None):
        # X: numpy array or (pinned) CPU torch tensor, float64 (the reference's dtype) or
        # float32.  float64 is copied as is and cast on the device, so the host never touches
        # the 29 M samples per patient.
        if isinstance(X, torch.Tensor) and X.is_cuda:
            # already resident (e.g. a channel subset gathered on the device by
            # processing_utils.device_subsample): no upload
            assert X.dim() == 3, 'features must be (trials, time, channels)'
            self.N, self.T, self.C = (int(v) for v in X.shape)
            # the producer (device_subsample.resident / gather_channels, or the caller's own
            # kernels) ran on another stream than this engine's lane: wait for it
            cur = torch.cuda.current_stream(X.device)
            ev = getattr(X, '_cpsd_ready', None)
            if ev is not None:
                cur.wait_event(ev)
            else:
                cur.wait_stream(torch.cuda.default_stream(X.device))
            Xc = X.contiguous()
            if Xc.dtype == torch.float64:
                self.X = ctx.empty((self.N * self.T, self.C))
                ctx.call('cpsd_cast_f64_f32', ptr(Xc), ptr(self.X), Xc.numel())
            else:
                self.X = Xc.to(torch.float32).view(self.N * self.T, self.C)
            self._host = Xc
            self.y = np.asarray(y).astype(np.int64)
            self.cls = np.asarray(cls_ids, dtype=np.int32)
            self.h2d_bytes = 0
            return
        if isinstance(X, torch.Tensor):
            host = X if X.is_contiguous() else X.contiguous()
        else:
            host = torch.from_numpy(np.ascontiguousarray(X))
        assert host.dim() == 3, 'features must be (trials, time, channels)'
        if host.dtype not in (torch.float32, torch.float64):
            host = host.to(torch.float64)
        self.N, self.T, self.C = (int(v) for v in host.shape)
        if not host.is_pinned():
            host = host.pin_memory()
        if upload_stream is not None:
            # all uploads of a streamed run go through ONE stream, in submission order: the first
            # job's data is complete (and its kernels start) while the later jobs still copy --
            # uploads issued on every job's own stream share the link and all finish late
            # (buffers come from the lane's own allocator pool; the upload stream only borrows them
            # between two events, so no cross-stream allocator bookkeeping is needed)
            lane = torch.cuda.current_stream(ctx.device)
            raw = torch.empty(host.shape, dtype=host.dtype, device=ctx.device)
            f64 = raw.dtype == torch.float64
            self.X = ctx.empty((self.N * self.T, self.C)) if f64 else raw.view(self.N * self.T, self.C)
            ev0 = torch.cuda.Event()
            ev0.record(lane)
            upload_stream.wait_event(ev0)
            with torch.cuda.stream(upload_stream):    # DMA only: kernels would queue behind the busy SMs
                raw.copy_(host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(upload_stream)
            lane.wait_event(ev)
            if f64:
                ctx.call('cpsd_cast_f64_f32', ptr(raw), ptr(self.X), raw.numel())
        else:
            raw = host.to(ctx.device, non_blocking=True)
            if raw.dtype == torch.float64:
                self.X = ctx.empty((self.N * self.T, self.C))
                ctx.call('cpsd_cast_f64_f32', ptr(raw), ptr(self.X), raw.numel())
            else:
                self.X = raw.view(self.N * self.T, self.C)
        self._host = host                                     # keep staging alive until done
        self.y = np.asarray(y).astype(np.int64)
        self.cls = np.asarray(cls_ids, dtype=np.int32)        # alignment class id per trial
        self.h2d_bytes = host.numel() * host.element_size()


class _TopkRetry(Exception):
    """A speculatively used top-k round failed its acceptance test: run the batch again."""


class _Batch(list):
    """The folds of one engine batch plus what travels with them (bagging seeds, replica ids)."""
    bag_seeds = None
    rep = None


class CVEngine:
    """See module docstring.  ``method``: 'mcca' | 'cca' | 'none'."""

    def __init__(self, target, cross, method='mcca', n_comp=None, regs=0.5, pca_var=0.8,
                 decoder_var=0.8, C=1.0, tar_in_train=True, device=None, max_batch=32,
                 dcd_epochs=0, max_newton=400, tol_newton=1e-9, tol_dcd=1e-4,
                 eig_sweeps=12, eig_tol=3e-7, use_tensor_cores=False, pool_solver='auto',
                 topk_block=128, topk_iters=8, topk_tol=5e-6, topk_rounds=3, n_lanes=2, lane=0,
                 topk_tf32_iters=5, topk_gap_tol=0.05, decoder='linear', class_weight=None, svc_tol=1e-3,
                 svc_gamma='scale', svc_max_iter=1000000, joint_cap=None, n_estimators=10,
                 replicas=None, upload_stream=None, speculative_topk=True):
        # decoder: 'linear' = one-vs-rest squared-hinge linear SVM (the north star's dual-CD
        # decoder); 'svc_rbf' / 'svc_linear' = libsvm-style C-SVC with one-vs-one votes, the
        # reference scripts' literal SVC(kernel=..., class_weight=...)
        # 'bag_svc_linear' / 'bag_svc_rbf': sklearn BaggingClassifier(estimator=SVC(kernel=...),
        # n_estimators) around the C-SVC -- scripts/aligned_decode_svm.py:262-265 -- as n_estimators
        # extra C-SVC problems per fold
        assert decoder in ('linear', 'svc_rbf', 'svc_linear', 'bag_svc_linear', 'bag_svc_rbf')
        assert class_weight in (None, 'balanced')
        self.decoder = decoder
        self.class_weight = class_weight
        self.svc_tol = float(svc_tol)
        self.svc_gamma = -1.0 if svc_gamma == 'scale' else float(svc_gamma)
        self.svc_max_iter = int(svc_max_iter)
        self.topk_gap_tol = float(topk_gap_tol)
        # replicas: further (target, cross) data sets with the SAME shapes and labels as the primary
        # one (independent jobs, e.g. the CV iterations of a streamed run, each with its own
        # upload): their folds share this engine's batches -- run(folds, rep=...) says which
        # replica a fold reads -- so the latency-bound solver launches are amortised over all of
        # them.  MCCA with the tensor-core projection only.
        self._replicas = list(replicas or [])
        self._upload_stream = upload_stream
        # decoder-PCA top-k solver: launch the rest of the batch on the first round's result and
        # verify its acceptance test with the batch's final read-back (one host sync per batch
        # instead of two; a batch whose first round is not accepted is run again the slow way)
        self.speculative_topk = bool(speculative_topk) and os.environ.get('CPSD_SPECULATIVE_TOPK', '1') != '0'
        # warm-started tile eigen-solves (CPSD_WARM_START=0: every solve starts from the identity)
        self.warm_start = os.environ.get('CPSD_WARM_START', '1') != '0'
        self.side_streams = os.environ.get('CPSD_SIDE_STREAMS', '1') != '0'
        self._runs_done = 0
        self._spec = False
        base = Context.get(device)
        self.lane = int(lane)
        self.stream = _lane_stream(base.device, self.lane)
        self.lane_idx = self.lane
        self.ctx = LaneContext(base, self.stream)       # every launch of this engine goes to its lane
        with torch.cuda.stream(self.stream):     # uploads + fold-invariant work on the lane's stream
            self._init(target, cross, method, n_comp, regs, pca_var, decoder_var, C, tar_in_train,
                       max_batch, dcd_epochs, max_newton, tol_newton, tol_dcd, eig_sweeps, eig_tol,
                       use_tensor_cores, pool_solver, topk_block, topk_iters, topk_tol, topk_rounds,
                       n_lanes)
            self.topk_tf32_iters = int(topk_tf32_iters)
        if joint_cap is not None:
            self._joint_qcap = int(joint_cap)
        self.n_estimators = int(n_estimators)

    def _init(self, target, cross, method, n_comp, regs, pca_var, decoder_var, C, tar_in_train,
              max_batch, dcd_epochs, max_newton, tol_newton, tol_dcd, eig_sweeps, eig_tol,
              use_tensor_cores, pool_solver, topk_block, topk_iters, topk_tol, topk_rounds, n_lanes):
        self.method = method
        if n_comp is None:
            n_comp = 30 if method == 'mcca' else (40 if method == 'jointpca' else 0.9)
        self.n_comp = n_comp
        self.regs = regs
        self.pca_var = pca_var
        self.decoder_var = decoder_var
        self.Csvm = float(C)
        self.tar_in_train = tar_in_train
        self.max_batch = max_batch
        self.dcd_epochs = dcd_epochs
        self.max_newton = max_newton
        self.tol_newton = tol_newton
        self.tol_dcd = tol_dcd
        self.eig_sweeps = eig_sweeps
        self.eig_tol = eig_tol
        self.use_tc = use_tensor_cores
        # (topk_tol: Ritz residual / theta_1 of every retained pair.  5e-6 is the level the fp32
        # Gram itself is accurate to; a looser bound lets the retained subspace drift by
        # resid / (gap at the cut), which flips near-tie labels on flat noisy spectra.)
        # decoder-stage PCA eigen-solver: 'topk' = block subspace iteration for the leading
        # components (falls back to the full solver when the requested variance is not reached
        # inside the block), 'full' = block Jacobi of the whole Gram, 'auto' = topk when the
        # pooled matrix is large enough for it to pay (at least 3 blocks of rows and columns)
        assert pool_solver in ('auto', 'topk', 'full')
        self.pool_solver = pool_solver
        self.topk_block = int(topk_block)
        self.topk_iters = int(topk_iters)
        self.topk_tol = float(topk_tol)
        self.topk_rounds = int(topk_rounds)
        # JointPCA(n_components=0.9) -- what the script's set_params hands to the class
        # (scripts/aligned_decode_svm_ncv.py:186-190): a variance fraction, i.e. a component count
        # that depends on the fold.  The batch is laid out for `_joint_qcap` columns (sized from a
        # fit on all trials, see _ensure_ready) and every fold's read-in matrices are cut at its own
        # count; the surplus columns of the pooled matrix are zero and change nothing downstream.
        self.joint_var = (method == 'jointpca' and isinstance(n_comp, (float, np.floating))
                          and 0 < n_comp < 1)
        self._joint_qcap = None
        if method == 'mcca' or (method == 'jointpca' and not self.joint_var):
            assert isinstance(n_comp, (int, np.integer)) and n_comp >= 1
        views = [target] + list(cross)
        ids, self.vocab = class_ids([v[2] if v[2] is not None else v[1] for v in views])
        us = self._upload_stream
        self.views = [View(self.ctx, v[0], v[1], v[2], i, us) for v, i in zip(views, ids)]
        self.rviews = [self.views]
        for rt, rc in self._replicas:
            rv = [rt] + list(rc)
            assert method == 'mcca' and len(rv) == len(views), 'replicas: MCCA, same patient count'
            vs_r = [View(self.ctx, v[0], v[1], v[2], i, us) for v, i in zip(rv, ids)]
            for a, b in zip(vs_r, self.views):
                assert (a.N, a.T, a.C) == (b.N, b.T, b.C), 'replicas must have the primary job\'s shapes'
            self.rviews.append(vs_r)
        self.J = len(self.rviews)
        self._replicas = None                      # the host arrays are not kept alive by the engine
        self.P = len(self.views)
        self.T = self.views[0].T
        assert all(v.T == self.T for v in self.views), 'all patients need the same time axis'
        self.Cmax = max(v.C for v in self.views)
        self.classes = np.unique(np.concatenate([v.y for v in self.views])).astype(np.int32)
        self.classes_dev = self.ctx.upload(self.classes, np.int32)
        self.packA = HostPack(self.ctx)
        self.packB = HostPack(self.ctx)
        self.packM = [HostPack(self.ctx), HostPack(self.ctx)]   # MCCA batches alternate
        self._pack_i = 0
        self.n_lanes = int(n_lanes)
        self._ws = {}
        self._sched = {}
        self.stats = {}
        self.profile = False
        self._marks = []
        self._prepare_cross()

    # ------------------------------------------------------------------ stage timing
    def mark(self, stage):
        """Records a CUDA event on the launching stream when profiling is on; the time between
        consecutive marks is attributed to the earlier mark's stage name."""
        if getattr(self, 'profile', False):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.ctx.device))
            self._marks.append((stage, ev))

    def collect_marks(self):
        """Returns {stage: milliseconds} accumulated since the last call (synchronises)."""
        out = {}
        marks = getattr(self, '_marks', [])
        if marks:
            torch.cuda.synchronize(self.ctx.device)
            for (name, e0), (_, e1) in zip(marks[:-1], marks[1:]):
                if name != 'end':
                    out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        self._marks = []
        return out

    # ------------------------------------------------------------------ workspace
    def ws(self, name, shape, dtype=F32):
        """Named persistent workspace tensor, grown on demand."""
        n = int(np.prod(shape))
        t = self._ws.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = self.ctx.empty((max(n, 1),), dtype)
            self._ws[name] = t
        return t[:n].view(*shape) if n else t[:0]

    def schedule(self, n_pad):
        if n_pad not in self._sched:
            nb = n_pad // 64
            host = np.zeros((nb - 1) * (nb // 2) * 2, dtype=np.int32)
            _lib.check(self.ctx.lib.cpsd_bj_schedule(n_pad, host.ctypes.data), 'bj_schedule')
            self._sched[n_pad] = self.ctx.upload(host, np.int32)
        return self._sched[n_pad]

    # ------------------------------------------------------------------ eigen helpers
    def eig_small(self, A, n_dev, n_fixed, nprob, lda, evals, evecs, ldv):
        """A: (nprob, lda, lda) tensor, n <= 128.  Sorted descending."""
        self.ctx.call('cpsd_eig_sym_small', ptr(A), lda, lda * lda, _p(n_dev), n_fixed, nprob,
                      ptr(evals), evals.shape[-1], ptr(evecs), ldv, ldv * ldv if evecs is not None
                      else 0, self.eig_sweeps + 3, self.eig_tol, ptr(None))

    def eig_block(self, K, n_pad, n_dev, n_fixed, nprob, evals, perm, tag):
        """K: (nprob, n_pad, n_pad), destroyed.  Eigenvalues (descending) + diagonal positions;
        eigenvectors come from eig_vecs() (rotation-log replay)."""
        nlog = int(self.ctx.lib.cpsd_bj_rlog_elems(n_pad, nprob, self.eig_sweeps))
        R = self.ws(tag + '_Rlog', (nlog,))
        fw = self.ws(tag + '_fw', (18 * nprob,))
        iw = self.ws(tag + '_iw', (2 * nprob,), I32)
        RT = self.ws(tag + '_RT', (nprob * (n_pad // 128) * 128 * 128,)) if self.use_tc else None
        self.ctx.call('cpsd_eig_sym_block', ptr(K), n_pad, n_pad * n_pad, n_pad, _p(n_dev),
                      n_fixed, nprob, ptr(self.schedule(n_pad)), ptr(R), ptr(fw), ptr(iw),
                      ptr(evals), ptr(perm), evals.shape[-1], self.eig_sweeps, self.eig_tol,
                      ptr(RT))
        return iw

    def eig_vecs(self, tag, n_pad, nprob, perm, k_dev, k_fixed, k_launch, E):
        """Leading eigenvectors (sorted columns) of the last eig_block(tag) call into E
        (nprob, n_pad, n_pad)."""
        R = self._ws[tag + '_Rlog']
        iw = self._ws[tag + '_iw']
        self.ctx.call('cpsd_bj_eigvecs', ptr(R), n_pad, nprob, ptr(self.schedule(n_pad)), ptr(iw),
                      ptr(perm), n_pad, _p(k_dev), k_fixed, k_launch, ptr(E), n_pad, n_pad * n_pad,
                      self.eig_sweeps)

    def eig_any(self, A, n_pad, n_dev, n_fixed, nprob, tag, ncols=None, vecs=True, out=None):
        """Sorted eigen-decomposition for any n_pad (A: (nprob, n_pad, n_pad), destroyed).
        A float64 tensor (n_pad <= 128 only) selects the fp64-matrix solver.
        Returns (evals (nprob, n_pad), evecs (nprob, n_pad, n_pad) with sorted columns; for
        n_pad > 128 only the leading ``ncols`` columns are computed)."""
        if out is not None:      # caller-owned (nprob, n_pad) / (nprob, n_pad, n_pad) outputs
            evals, evecs = out
        else:
            evals = self.ws(tag + '_ev', (nprob, n_pad))
            evecs = self.ws(tag + '_evec', (nprob, n_pad, n_pad))
        if nprob == 0:
            return evals, evecs
        if A.dtype == torch.float64:
            if n_pad <= 128:
                self.ctx.call('cpsd_eig_sym_small_f64', ptr(A), n_pad, n_pad * n_pad, _p(n_dev),
                              n_fixed, nprob, ptr(evals), n_pad, ptr(evecs if vecs else None), n_pad,
                              n_pad * n_pad, self.eig_sweeps + 6, 1e-10, ptr(None))
                return evals, evecs
            # 128 < n <= 256 (patients with more than 128 electrodes): fp64 one-sided Jacobi on an
            # L2-resident workspace (csrc/solve64.cu)
            assert n_pad <= 256
            w64 = self.ws(tag + '_e64', (int(self.ctx.lib.cpsd_eig_sym_f64_ws_elems(nprob, n_pad)),),
                          torch.float64)
            if vecs:
                evecs.zero_()
            self.ctx.call('cpsd_eig_sym_f64', ptr(A), n_pad, n_pad * n_pad, _p(n_dev), n_fixed, nprob,
                          ptr(evals), n_pad, ptr(evecs if vecs else None), n_pad, n_pad * n_pad, 40,
                          ptr(w64), n_pad, ptr(None))
            return evals, evecs
        if n_pad <= 128:
            self.eig_small(A, n_dev, n_fixed, nprob, n_pad, evals, evecs, n_pad)
            return evals, evecs
        perm = self.ws(tag + '_perm', (nprob, n_pad), I32)
        self.eig_block(A, n_pad, n_dev, n_fixed, nprob, evals, perm, tag)
        nc = n_pad if ncols is None else ncols
        self.eig_vecs(tag, n_pad, nprob, perm, ptr(None), nc, nc, evecs)
        return evals, evecs

    # ------------------------------------------------------------------ warm-started tile solves
    def eig_warm(self, A, n_dev, n_fixed, nsel, evals, evecs, sel=None, out_idx=None, V0=None,
                 v0_idx=None):
        """fp64 tile solve (n <= 128) of the problems ``sel`` (device int list, None = 0..nsel-1) of
        A (nprob, 128, 128); results go to slot ``out_idx[problem]`` of evals (.., 128) / evecs
        (.., 128, 128); the eigenvector accumulator starts from ``V0[v0_idx[problem]]`` (the caller
        has rotated A into that basis, see ``rotate_sym``)."""
        if nsel == 0:
            return
        self.ctx.call('cpsd_eig_sym_small_f64_warm', ptr(A), 128, 128 * 128, _p(n_dev), n_fixed,
                      _p(sel), nsel, _p(out_idx), ptr(evals), evals.shape[-1], ptr(evecs), 128,
                      128 * 128, ptr(V0), 128, 128 * 128, _p(v0_idx), self.eig_sweeps + 6, 1e-10,
                      ptr(None))

    def rotate_sym(self, A, Q, nsel, tag, sel=None, base=None):
        """A[sel[i]] <- Q[base[i]]^T A[sel[i]] Q[base[i]] in fp64 (A, Q: (.., 128, 128); sel / base:
        device int lists, None = i)."""
        if nsel == 0:
            return
        W = self.ws(tag + '_rotw', (nsel, 128, 128), torch.float64)
        st = 128 * 128
        self.ctx.call('cpsd_dgemm_batched', 0, 128, 128, 128, 1.0, ptr(A), 128, st, _p(sel), ptr(Q), 128,
                      st, _p(base), 0.0, ptr(None), 128, 0, _p(None), ptr(W), 128, st, _p(None), nsel)
        self.ctx.call('cpsd_dgemm_batched', 1, 128, 128, 128, 1.0, ptr(Q), 128, st, _p(base), ptr(W), 128,
                      st, _p(None), 0.0, ptr(None), 128, 0, _p(None), ptr(A), 128, st, _p(sel), nsel)

    def ortho_bases(self, evecs, slots, nb, Q, Qf, b0):
        """Q[b0 + i] = fp64 re-orthonormalised copy of the fp32 eigenvector matrix evecs[slots[i]]
        (one Newton-Schulz step Q (3 I - Q^T Q) / 2: orthogonal to ~1e-12, so that a rotation by it
        is an exact similarity), Qf its fp32 rounding (the solver's starting accumulator)."""
        if nb == 0:
            return
        st = 128 * 128
        Qd = self.ws('ob_qd', (nb, 128, 128), torch.float64)
        G = self.ws('ob_g', (nb, 128, 128), torch.float64)
        self.ctx.call('cpsd_cast_f32_f64_idx', ptr(evecs), st, _p(slots), ptr(Qd), st, st, nb)
        self.ctx.call('cpsd_dgemm_batched', 1, 128, 128, 128, 1.0, ptr(Qd), 128, st, _p(None), ptr(Qd),
                      128, st, _p(None), 0.0, ptr(None), 128, 0, _p(None), ptr(G), 128, st, _p(None), nb)
        self.ctx.call('cpsd_dgemm_batched', 0, 128, 128, 128, -0.5, ptr(Qd), 128, st, _p(None), ptr(G),
                      128, st, _p(None), 1.5, ptr(Qd), 128, st, _p(None), ptr(Q, b0 * st), 128, st,
                      _p(None), nb)
        self.ctx.call('cpsd_cast_f64_f32', ptr(Q, b0 * st), ptr(Qf, b0 * st), nb * st)

    def eig_topk(self, K, n_pad, n_dev, nprob, m, evals, tag, thr, mode, kcap, k2, tol=None,
                 gap_tol=None):
        """Leading eigen-pairs of K (nprob, n_pad, n_pad) by subspace iteration; K keeps its
        leading block.  Fills evals / k2 and returns (V tensor view, ldv, strideV) when every
        problem reached its component count inside the block with converged Ritz pairs, else
        None (the caller falls back to the full solver).  Generator: yields before every
        blocking read-back so that another lane's work can be queued first."""
        ctx = self.ctx
        nws = int(ctx.lib.cpsd_eig_topk_ws_elems(n_pad, m, nprob))
        ws = self.ws(tag + '_tkws', (nws,))
        tot = self.ws(tag + '_tktot', (nprob,))
        resid = self.ws(tag + '_tkres', (nprob, m))
        status = self.ws(tag + '_tkst', (nprob,), I32)
        voff = int(ctx.lib.cpsd_eig_topk_voff(n_pad, m, nprob))
        V = ws[voff:voff + nprob * 2 * n_pad * m]
        info = {'rounds': 0, 'ok': False, 'tag': tag}
        self.stats['topk'] = info
        self.stats.setdefault('topk_log', []).append(info)
        del self.stats['topk_log'][:-16]
        tc = self.use_tc and m == 128 and n_pad % 128 == 0
        if tc:
            # K Q on the tensor cores: hi/lo workspace + tensor maps (re-encoded only when the
            # Gram or the workspace moved)
            tcw = self.ws(tag + '_tkc', (int(ctx.lib.cpsd_topk_tc_ws_elems(n_pad, nprob)),))
            nb = int(ctx.lib.cpsd_topk_tc_map_bytes(nprob))
            maps = self.ws(tag + '_tkm', (nb + 64,), torch.uint8)
            mp = (maps.data_ptr() + 63) & ~63
            key = (K.data_ptr(), tcw.data_ptr(), mp, n_pad, nprob)
            cache = getattr(self, '_tkc_key', None)
            if not isinstance(cache, dict):
                cache = self._tkc_key = {}
            if cache.get(tag, (None, None))[0] != key:
                stage = torch.empty((nb + 64,), dtype=torch.uint8).pin_memory()
                ctx.call('cpsd_topk_tc_encode', ptr(K), n_pad, n_pad * n_pad, n_pad, nprob, ptr(tcw),
                         ctypes.c_void_p(mp), ctypes.c_void_p(stage.data_ptr()))
                cache[tag] = (key, stage)
        tol = self.topk_tol if tol is None else tol
        gtol_ = self.topk_gap_tol if gap_tol is None else gap_tol
        if self._spec and tag == 'pool':
            f64s = getattr(self, '_tk_f64', None) or {}
            self._topk_round(tc, K, n_pad, n_dev, nprob, m, 1, ws, evals, tot, resid, status,
                             tcw if tc else None, mp if tc else None, int(f64s.get(tag, False)))
            ctx.call('cpsd_select_k_total', ptr(evals), evals.shape[-1], ptr(None), m, ptr(tot),
                     thr, mode, 1, min(kcap, m), ptr(k2), 1, nprob)
            info.update(rounds=1, speculative=True, f64_gram=bool(f64s.get(tag, False)))
            self._spec_check = dict(k2=k2, resid=resid, evals=evals, status=status, m=m, mode=mode,
                                    tol=tol, gtol=gtol_, nprob=nprob, info=info)
            self._k2_max = max(m - 8, 1)          # upper bound: sizes shared memory only
            return V, m, 2 * n_pad * m
        prev = None
        # Gram of the Cholesky-QR steps: fp32 first; a tag whose block once lost rank that way
        # (leading spectrum spanning > ~3e3) uses the fp64-accumulated Gram from then on
        f64 = getattr(self, '_tk_f64', None)
        if f64 is None:
            f64 = self._tk_f64 = {}
        # up to topk_rounds rounds; more (at most twice as many) only while every extra round
        # still shrinks the worst residual 4x -- far cheaper than the full solver it avoids
        rnd, fresh = 0, 1
        while rnd < 2 * self.topk_rounds:
            self._topk_round(tc, K, n_pad, n_dev, nprob, m, fresh, ws, evals, tot, resid, status,
                             tcw if tc else None, mp if tc else None, int(f64.get(tag, False)))
            ctx.call('cpsd_select_k_total', ptr(evals), evals.shape[-1], ptr(None), m, ptr(tot),
                     thr, mode, 1, min(kcap, m), ptr(k2), 1, nprob)
            yield 'sync'
            k2h = k2.cpu().numpy()
            rh = resid.cpu().numpy()
            ev0 = evals[:, 0].cpu().numpy()
            rnd += 1
            fresh = 0
            info['rounds'] = rnd
            info['k2_max'] = int(k2h.max())
            info['f64_gram'] = bool(f64.get(tag, False))
            if status.cpu().numpy().any():
                if not f64.get(tag, False):
                    f64[tag] = True            # start over with the fp64 Gram
                    rnd, fresh, prev = 0, 1, None
                    continue
                info['why'] = 'rank-deficient block'
                return None
            if mode == 0 and (k2h >= m - 8).any():
                info['why'] = 'variance threshold not crossed inside the block'
                return None
            worst = max(float(rh[f, :k2h[f]].max()) / max(float(ev0[f]), 1e-30)
                        for f in range(nprob))
            info['resid'] = worst
            # the retained subspace is only as good as residual / (gap at the cut): with a
            # near-degenerate cut (flat noisy spectra) the block result is left to the full solver
            evh = evals[:, :m].cpu().numpy()
            kk = np.clip(k2h, 1, m - 1)
            ar = np.arange(nprob)
            gap = evh[ar, kk - 1] - evh[ar, kk]
            rmax = np.array([rh[f, :k2h[f]].max() if k2h[f] > 0 else 0.0 for f in range(nprob)])
            gtol = self.topk_gap_tol if gap_tol is None else gap_tol
            gap_ok = bool((rmax <= gtol * np.maximum(gap, 0.0)).all()) or not np.isfinite(gtol)
            info['gap_ok'] = gap_ok
            if worst <= tol and gap_ok:
                info['ok'] = True
                self._k2_max = max(int(k2h.max()), 1)
                return V, m, 2 * n_pad * m
            if rnd >= self.topk_rounds and (prev is None or worst > 0.25 * prev):
                info['why'] = 'residual stalled'
                return None
            prev = worst
        return None

    def _topk_round(self, tc, K, n_pad, n_dev, nprob, m, fresh, ws, evals, tot, resid, status, tcw,
                    mp, f64_gram):
        """One round (topk_iters subspace iterations + Rayleigh-Ritz + residuals) of the top-k solver."""
        ctx = self.ctx
        if tc:
            ctx.call('cpsd_eig_sym_topk_tc', ptr(K), n_pad, n_pad * n_pad, n_pad, _p(n_dev), 0,
                     nprob, m, self.topk_iters, fresh, ptr(ws), ptr(evals),
                     evals.shape[-1], ptr(tot), ptr(resid), ptr(status), self.eig_sweeps + 3,
                     self.eig_tol, ptr(tcw), ctypes.c_void_p(mp), self.topk_tf32_iters, f64_gram)
        else:
            ctx.call('cpsd_eig_sym_topk', ptr(K), n_pad, n_pad * n_pad, n_pad, _p(n_dev), 0,
                     nprob, m, self.topk_iters, fresh, ptr(ws), ptr(evals),
                     evals.shape[-1], ptr(tot), ptr(resid), ptr(status), self.eig_sweeps + 3,
                     self.eig_tol, f64_gram)

    def _verify_spec(self):
        """Acceptance test of a speculatively used top-k round (same criteria as eig_topk), run on
        the batch's final read-back.  Raises _TopkRetry when the round would not have been accepted."""
        chk = self.__dict__.pop('_spec_check', None)
        if chk is None:
            return
        m, nprob, info = chk['m'], chk['nprob'], chk['info']
        k2h = chk['k2'].cpu().numpy()
        rh = chk['resid'].cpu().numpy()
        evh = chk['evals'][:, :m].cpu().numpy()
        ok = not chk['status'].cpu().numpy().any()
        if ok and chk['mode'] == 0 and (k2h >= m - 8).any():
            ok = False
        if ok:
            rmax = np.array([rh[f, :k2h[f]].max() if k2h[f] > 0 else 0.0 for f in range(nprob)])
            worst = float((rmax / np.maximum(evh[:, 0], 1e-30)).max())
            kk = np.clip(k2h, 1, m - 1)
            ar = np.arange(nprob)
            gap = evh[ar, kk - 1] - evh[ar, kk]
            gap_ok = bool((rmax <= chk['gtol'] * np.maximum(gap, 0.0)).all()) or not np.isfinite(chk['gtol'])
            info.update(resid=worst, gap_ok=gap_ok, k2_max=int(k2h.max()))
            ok = worst <= chk['tol'] and gap_ok
        info['ok'] = ok
        if not ok:
            raise _TopkRetry()
        self._k2_max = max(int(k2h.max()), 1)

    # ------------------------------------------------------------------ tensor-core projection
    def _tc_proj_ready(self, Q):
        """Tensor-core pooled projection (csrc/tc_proj.cu): any channel count <= 256 (two resident
        panels above 128), <= 128 latent columns (chunks of 32) and <= 16 patients.  Splits every
        patient into tf32 hi / lo once (row stride padded to a multiple of 4 floats, which is what
        lets odd channel counts through TMA) and encodes the TMA tensor maps."""
        if not self.use_tc or Q > 128 or self.P * self.J > 128:
            return False
        if any(v.C > 256 for v in self.views):
            return False
        if getattr(self, '_tcp', None) is None:
            ctx = self.ctx
            allv = [vw for rv in self.rviews for vw in rv]        # (replica, patient) order
            maps_h = torch.zeros((2 * len(allv), 128), dtype=torch.uint8).pin_memory()
            keep = []
            for i, vw in enumerate(allv):
                ldd = _ceil(vw.C, 4)
                rows = vw.N * vw.T
                hi, lo = ctx.empty((rows, ldd)), ctx.empty((rows, ldd))
                ctx.call('cpsd_split_tf32_2d', ptr(vw.X), vw.C, rows, vw.C, ptr(hi), ptr(lo), ldd)
                for u, t in enumerate((hi, lo)):
                    _lib.check(ctx.lib.cpsd_tmap_encode_f32(
                        ctypes.c_void_p(maps_h[2 * i + u].data_ptr()), ptr(t), rows, vw.C, ldd, 128),
                        'tmap_encode')
                keep += [hi, lo]
            self._tcp = dict(xmaps=maps_h.to(ctx.device, non_blocking=True), xmaps_host=maps_h,
                             split=keep, cap=0, nq=0, ltc=128 if self.Cmax <= 128 else 256,
                             ntr=np.array([v.N for v in allv], dtype=np.int32),
                             nch=np.array([v.C for v in allv], dtype=np.int32),
                             sms=torch.cuda.get_device_properties(ctx.device).multi_processor_count)
        return True

    def _tc_proj_ws(self, nprob, Q):
        """L^T hi / lo (nprob * nq, 32, ltc), mu L (nprob * nq, 32) and their tensor maps."""
        tcp = self._tcp
        nq = -(-Q // 32)
        if tcp['cap'] < nprob or tcp['nq'] != nq:
            ctx = self.ctx
            cap = max(nprob, self.max_batch * self.P)
            ltc = tcp['ltc']
            tcp['lthi'] = ctx.zeros((cap * nq, 32, ltc))
            tcp['ltlo'] = ctx.zeros((cap * nq, 32, ltc))
            tcp['mul'] = ctx.zeros((cap * nq, 32))
            mh = torch.zeros((2, 128), dtype=torch.uint8).pin_memory()
            for u, t in enumerate((tcp['lthi'], tcp['ltlo'])):
                _lib.check(ctx.lib.cpsd_tmap_encode_f32(ctypes.c_void_p(mh[u].data_ptr()), ptr(t),
                                                        cap * nq * 32, ltc, ltc, 32), 'tmap_encode')
            tcp['ltmaps'] = mh.to(ctx.device, non_blocking=True)
            tcp['ltmaps_host'] = mh               # pinned source stays alive until the copy ran
            tcp['cap'], tcp['nq'] = cap, nq
        return tcp

    def _view_slots(self, B, n_pad, Cm):
        """Per-view statistics of the MCCA fit (mean, spectrum and eigenvectors of the centred
        condition-average scatter) live in slots: [0, res) are the per-fold target problems of
        the current batch, the rest caches the cross patients' problems keyed by (view, shared
        class set) -- they do not depend on the fold (AlignMCCA.py:140-154 recomputes them for
        every fold; here they are solved once and reused)."""
        vs = getattr(self, '_vs', None)
        if vs is None or vs['res'] < B or vs['npad'] != n_pad:
            res = max(B, self.max_batch)
            cap = res + max(1024, 2 * res * max(self.P - 1, 1))
            vs = dict(res=res, cap=cap, npad=n_pad, keys={}, next=res,
                      mu=self.ctx.zeros((cap, Cm)), ev=self.ctx.zeros((cap, n_pad)),
                      evec=self.ctx.zeros((cap, n_pad, n_pad)))
            self._vs = vs
        return vs

    def _sm_count(self):
        n = getattr(self, '_sms', None)
        if n is None:
            n = self._sms = torch.cuda.get_device_properties(self.ctx.device).multi_processor_count
        return n

    def _view_bases(self):
        """Warm-start bases of the per-view scatter eigenproblems: for every (replica, patient) the
        eigenvectors of the first problem solved for it, as fp64 (rotation) and fp32 (starting
        accumulator) matrices; ``tab[replica * P + patient]`` = base index or -1."""
        vb = getattr(self, '_vb', None)
        if vb is None:
            cap = self.J * self.P
            vb = self._vb = dict(cap=cap, n=0, tab=-np.ones(cap, dtype=np.int32),
                                 Q=self.ctx.zeros((cap, 128, 128), torch.float64),
                                 Qf=self.ctx.zeros((cap, 128, 128)))
        return vb

    def _tg_rotate(self):
        """Second use of a resident engine: moves the target's per-trial Grams (and column sums)
        into the eigenbasis of their all-trials sum, once.  Every fold's train-set Gram / covariance
        assembled from them is then nearly diagonal: the signal-rank solve needs the eigenvalues
        only, the PCA solve of the CCA path starts its eigenvector accumulator from that basis."""
        tg = getattr(self, 'tg', None)
        if tg is None or tg.get('rot') or not self.warm_start:
            return
        ctx, J, N = self.ctx, self.J, self.views[0].N
        st = 128 * 128
        G = tg['trial']
        nrow = J * N + J
        A = (-G[J * N:]).contiguous()                       # the all-trials Grams (stored negated)
        ev = ctx.empty((J, 128))
        evec = ctx.empty((J, 128, 128))
        self.eig_warm(A, ptr(None), self.views[0].C, J, ev, evec)
        Q = ctx.zeros((J, 128, 128), torch.float64)
        Qf = ctx.zeros((J, 128, 128))
        self.ortho_bases(evec, None, J, Q, Qf, 0)
        base = np.concatenate([np.repeat(np.arange(J), N), np.arange(J)]).astype(np.int32)
        base_dev = ctx.upload(base)
        self.rotate_sym(G, Q, nrow, 'tg', base=base_dev)
        if 'sums' in tg:
            S = tg['sums']
            tg['sums0'] = S.clone()
            ctx.call('cpsd_dgemm_batched', 0, 1, 128, 128, 1.0, ptr(tg['sums0']), 128, 128, _p(None),
                     ptr(Q), 128, st, ptr(base_dev), 0.0, ptr(None), 128, 0, _p(None), ptr(S), 128, 128,
                     _p(None), nrow)
        tg.update(rot=True, Q=Q, Qf=Qf, base_dev=base_dev)

    def _xcache(self, R, XR):
        """Reduced coordinates Zx (K*T x (P-1)R) of the cross patients' condition averages and
        their cross-scatter Gxx, one slot per shared class set (fold-invariant)."""
        xc = getattr(self, '_xc', None)
        if xc is None or xc['R'] != R:
            cap = max(32, self.max_batch + 16)
            ncls = len(set.intersection(*self.cross_classes)) if self.P > 1 else 1
            KT = max(ncls, 1) * self.T
            xc = dict(R=R, cap=cap, KT=KT, keys={},
                      Zx=self.ctx.empty((cap, KT, max(XR, 1))),
                      Gxx=self.ctx.empty((cap, max(XR, 1), max(XR, 1))))
            self._xc = xc
        return xc

    def _slot_means(self, mu, slot, Cm):
        idx = torch.from_numpy(slot.ravel()).to(mu.device)
        return mu[idx].view(slot.shape[0], slot.shape[1], Cm).cpu().numpy()

    def gram_scatter(self, kernel, descs, nprob, p, q, out, tiles):
        """Launches a scatter-matrix Gram; the fp64 variant splits the row segments over
        several CTAs per tile when the batch alone cannot fill the GPU."""
        if kernel == 'cpsd_gram_tn_f64':
            nsplit = min(16, max(1, 592 // max(nprob * tiles, 1)))
            if nsplit > 1:
                ldo = out.shape[-1]
                part = self.ws('gram_part', (int(self.ctx.lib.cpsd_gram_tn_split_ws_elems(
                    nprob, p, ldo, nsplit)),), torch.float64)
                self.ctx.call('cpsd_gram_tn_f64_split', descs, nprob, p, q, ldo, nsplit, ptr(part))
                return
        self.ctx.call(kernel, descs, nprob, p, q)

    def scatter(self, name, nprob, n_pad):
        """Workspace + kernel name for a batch of scatter matrices that feed an eigen-solver:
        fp64 accumulation and the fp64-matrix solver whenever the tile solver applies."""
        if n_pad <= 256:
            return self.ws(name, (nprob, n_pad, n_pad), torch.float64), 'cpsd_gram_tn_f64'
        return self.ws(name, (nprob, n_pad, n_pad)), 'cpsd_gram_tn'

    # ------------------------------------------------------------------ fold-invariant work
    def _prepare_cross(self):
        """Class means of every cross patient (all classes), and for MCCA their signal ranks,
        for CCA / none their PCA bases -- none of it depends on the fold."""
        ctx, T = self.ctx, self.T
        pk = HostPack(ctx)
        self.cm = [None] * self.P          # class-mean tensors (n_cls_v * T, C_v)
        self.cm_row = [None] * self.P      # class id -> slot in cm (or -1)
        recs = np.zeros(self.P - 1, dtype=_lib.CLASS_MEAN_DESC)
        offs = []
        for v in range(1, self.P):
            vw = self.views[v]
            present = np.unique(vw.cls)
            row = -np.ones(len(self.vocab), dtype=np.int64)
            row[present] = np.arange(len(present))
            self.cm_row[v] = row
            order = np.argsort(vw.cls, kind='stable')
            counts = np.bincount(vw.cls, minlength=len(self.vocab))[present]
            mptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
            offs.append((pk.add_ints(mptr), pk.add_ints(order.astype(np.int32)), len(present)))
            self.cm[v] = ctx.empty((self.J, len(present) * T, vw.C))     # one slab per replica
        if self.P > 1:
            J = self.J
            recs = np.zeros((self.P - 1, J), dtype=_lib.CLASS_MEAN_DESC)
            pk.reserve_ints()
            for i, v in enumerate(range(1, self.P)):
                o_ptr, o_mem, ns = offs[i]
                for j in range(J):
                    vw = self.rviews[j][v]
                    recs[i, j] = (addr(vw.X), pk.iaddr(o_ptr), pk.iaddr(o_mem), addr(self.cm[v][j]), ns,
                                  T * vw.C, 0, 0)
            d_off = pk.add_descs(recs.ravel())
            pk.upload()
            for i, v in enumerate(range(1, self.P)):
                # one launch per view (TC differs between views), all replicas of the view in it
                ctx.call('cpsd_class_mean',
                         ctypes_off(pk.daddr(d_off), i * J * _lib.CLASS_MEAN_DESC.itemsize), J,
                         offs[i][2], T * self.views[v].C)
        self.cross_classes = [set(np.unique(self.views[v].cls).tolist())
                              for v in range(1, self.P)]
        if self.method == 'mcca':
            # the ranks stay on the device until the first batch needs them on the host
            # (_ensure_ready): constructing an engine does not block on its own uploads
            self._rank_dev = self._ranks_full(range(1, self.P))     # (replica, view) order
            self.cross_rank = None
            self._target_trial_grams()
            self._cross_class_stats()
            self._keep = [pk]                # staging of the kernels still in flight
        else:
            if self.method in ('cca', 'none') and self.Cmax <= 128 and self.J == 1:
                self._target_trial_grams(with_sums=True)       # per-fold PCA covariance by downdate
            if self.P > 1 and self.method != 'jointpca':
                self._cross_pca()
            torch.cuda.current_stream(self.ctx.device).synchronize()

    def _ensure_ready(self):
        """Blocking tail of the constructor: signal ranks of the cross patients -> host."""
        if self.method == 'mcca' and self.cross_rank is None:
            with torch.cuda.stream(self.stream):
                self.cross_rank = self._rank_dev.cpu().numpy().astype(np.int32).reshape(self.J, self.P - 1)
            self._keep = None
        if self.joint_var and self._joint_qcap is None:
            self._joint_qcap = 32
            tv = self.views[0]
            res = self._batch_mcca([(np.arange(tv.N), np.zeros(0, dtype=np.int64))], False,
                                   align_only=True)
            k_all = int(res['k'][0])
            self._joint_qcap = min(120, _ceil(k_all + max(4, k_all // 4), 4))

    def _cross_class_stats(self):
        """Per-class Gram matrices M_k^T M_k and column sums (fp64) of every cross patient's
        condition averages (M_k: the T rows of class k), for every replica.  The centred scatter of
        the averages restricted to ANY shared class set -- what AlignMCCA.py:140-154 feeds to the
        per-view SVD, one set per distinct loss of classes among the target's train trials -- is then
        sum_k G_k - s s^T / n with s = sum_k s_k: a list sum instead of a pass over the averages per
        (patient, class set).  Rows of the two tables: ccs_off[v] + replica * n_classes_v + class slot."""
        self.ccs = None
        if self.P < 2 or self.Cmax > 128 or not self.warm_start:
            return
        ctx, T, J = self.ctx, self.T, self.J
        ncl = [0] + [self.cm[v].shape[1] // T for v in range(1, self.P)]
        off = np.concatenate([[0, 0], np.cumsum([J * n for n in ncl[1:]])]).astype(np.int64)
        tot = int(off[-1])
        G = ctx.zeros((tot, 128, 128), torch.float64)
        S = ctx.zeros((tot, 128), torch.float64)
        pk = HostPack(ctx)
        o_zero = pk.add_ints([0])
        pk.reserve_ints()
        recs = np.zeros(tot, dtype=_lib.GRAM_TN_DESC)
        r = 0
        for v in range(1, self.P):
            C, n = self.views[v].C, ncl[v]
            for j in range(J):
                a = addr(self.cm[v][j]) + 4 * T * C * np.arange(n, dtype=np.int64)
                sl = slice(r, r + n)
                recs['A'][sl] = recs['B'][sl] = a
                recs['p'][sl] = recs['q'][sl] = recs['lda'][sl] = recs['ldb'][sl] = C
                r += n
            # (the replicas' slabs of a patient are contiguous: one launch per patient)
            ctx.call('cpsd_trial_colsum_f64', ptr(self.cm[v]), J * n, T, C, C, ptr(S, int(off[v]) * 128), 128)
        recs['segA'] = recs['segB'] = pk.iaddr(o_zero)
        recs['out'] = addr(G) + 8 * 128 * 128 * np.arange(tot, dtype=np.int64)
        recs['nseg'], recs['seg_len'] = 1, T
        recs['ldo'], recs['sym'], recs['alpha'] = 128, 1, 1.0
        d = pk.add_descs(recs)
        pk.upload()
        ctx.call('cpsd_gram_tn_f64', pk.daddr(d), tot, self.Cmax, self.Cmax)
        self.ccs = dict(G=G, S=S, off=off, ncl=ncl, keep=pk)

    def _ranks_full(self, vs):
        """AlignMCCA.n_components_var on all trials of the given views (AlignMCCA.py:146-150)."""
        vs = list(vs)
        if not (0 < self.pca_var < 1) or not vs:
            return torch.full((getattr(self, 'J', 1) * len(vs),), int(self.n_comp), dtype=I32,
                              device=self.ctx.device)
        ctx, T, Cm = self.ctx, self.T, self.Cmax
        n_pad = _ceil(Cm, 128) if Cm > 128 else 128
        pk = HostPack(ctx)
        J = getattr(self, 'J', 1)
        nprob = J * len(vs)                       # problem (j, i) = replica j, view vs[i]
        G, gram = self.scatter('rk_G', nprob, n_pad)
        G.zero_()
        seg = [pk.add_ints(np.arange(self.views[v].N, dtype=np.int32) * T) for v in vs]
        cdim = pk.add_ints([self.views[v].C for v in vs] * J)
        pk.reserve_ints()
        recs = np.zeros(nprob, dtype=_lib.GRAM_TN_DESC)
        for j in range(J):
            for i, v in enumerate(vs):
                vw = self.rviews[j][v]
                q = j * len(vs) + i
                recs[q] = (addr(vw.X), addr(vw.X), pk.iaddr(seg[i]), pk.iaddr(seg[i]), 0, 0,
                           addr(G, q * n_pad * n_pad), vw.N, T, vw.C, vw.C, vw.C, vw.C, n_pad, 1,
                           1.0, 0)
        d = pk.add_descs(recs)
        pk.upload()
        self.gram_scatter(gram, pk.daddr(d), nprob, Cm, Cm, G, 3)
        cd = ctypes_int_ptr(pk.iaddr(cdim))
        evals, _ = self.eig_any(G, n_pad, cd, 0, nprob, 'rk', vecs=False)
        k = self.ws('rk_k', (nprob,), I32)
        ctx.call('cpsd_select_k', ptr(evals), n_pad, cd, 0, float(self.pca_var), 1, 0, 1 << 30,
                 ptr(k), 1, nprob)
        self._keep_rk = pk
        return k.clone()

    def _target_trial_grams(self, with_sums=False):
        """Uncentred per-trial scatter matrices X_t^T X_t of the target (fp64), for every replica,
        followed by MINUS their sum per replica: the train-set Gram of a fold is
        -( -G_all + sum of the held-out trials' matrices ), one cpsd_sum_mats_f64 list per fold
        (rows: [replica * N + trial] ..., then [J * N + replica])."""
        self.tg = None
        tv = self.views[0]
        if tv.C > 128 or (not with_sums and not (0 < self.pca_var < 1)):
            return
        ctx, T, J, N = self.ctx, self.T, self.J, tv.N
        pk = HostPack(ctx)
        seg = pk.add_ints(np.arange(N, dtype=np.int32) * T)
        o_ptr = pk.add_ints(np.arange(J + 1, dtype=np.int32) * N)
        o_all = pk.add_ints(np.arange(J * N, dtype=np.int32))
        pk.reserve_ints()
        G = ctx.zeros((J * N + J, 128, 128), torch.float64)
        recs = np.zeros(J * N, dtype=_lib.GRAM_TN_DESC)
        tn = np.tile(np.arange(N, dtype=np.int64), J)
        recs['A'] = recs['B'] = np.repeat(np.array([addr(self.rviews[j][0].X) for j in range(J)],
                                                   dtype=np.int64), N)
        recs['segA'] = recs['segB'] = pk.iaddr(seg) + 4 * tn
        recs['out'] = addr(G) + 8 * 128 * 128 * np.arange(J * N, dtype=np.int64)
        recs['nseg'], recs['seg_len'] = 1, T
        recs['p'] = recs['q'] = recs['lda'] = recs['ldb'] = tv.C
        recs['ldo'], recs['sym'], recs['alpha'] = 128, 1, 1.0
        d = pk.add_descs(recs)
        pk.upload()
        ctx.call('cpsd_gram_tn_f64', pk.daddr(d), J * N, tv.C, tv.C)
        ctx.call('cpsd_sum_mats_f64', ptr(None), ptr(G), 128 * 128, ctypes_int_ptr(pk.iaddr(o_ptr)),
                 ctypes_int_ptr(pk.iaddr(o_all)), -1.0, ptr(G, J * N * 128 * 128), 128 * 128, 128 * 128, J)
        self.tg = dict(trial=G, keep=pk)
        if with_sums:
            # per-trial column sums (fp64), then MINUS their total, in the same row order as G
            S = ctx.zeros((J * N + J, 128), torch.float64)
            for j in range(J):
                ctx.call('cpsd_trial_colsum_f64', ptr(self.rviews[j][0].X), N, T, tv.C, tv.C,
                         ptr(S, j * N * 128), 128)
            ctx.call('cpsd_sum_mats_f64', ptr(None), ptr(S), 128, ctypes_int_ptr(pk.iaddr(o_ptr)),
                     ctypes_int_ptr(pk.iaddr(o_all)), -1.0, ptr(S, J * N * 128), 128, 128, J)
            self.tg['sums'] = S

    def _cross_pca(self):
        """sklearn PCA(n_comp) of every cross patient's (trials*time, channels) matrix
        (cross_pt_decoders.py:234-235; fold-invariant)."""
        ctx, T, Cm = self.ctx, self.T, self.Cmax
        vs = list(range(1, self.P))
        nv = len(vs)
        n_pad = _ceil(Cm, 128) if Cm > 128 else 128
        pk = HostPack(ctx)
        self.cross_mu = self.ws('cp_mu', (nv, Cm))
        cov, gram = self.scatter('cp_cov', nv, n_pad)
        cov.zero_()
        seg = [pk.add_ints(np.arange(self.views[v].N, dtype=np.int32) * T) for v in vs]
        cdim = pk.add_ints([self.views[v].C for v in vs])
        pk.reserve_ints()
        r1 = np.zeros(nv, dtype=_lib.COLSUM_DESC)
        r2 = np.zeros(nv, dtype=_lib.GRAM_TN_DESC)
        for i, v in enumerate(vs):
            vw = self.views[v]
            r1[i] = (addr(vw.X), pk.iaddr(seg[i]), addr(self.cross_mu, i * Cm), vw.N, T, vw.C,
                     vw.C, 1.0 / (vw.N * T), 0)
            r2[i] = (addr(vw.X), addr(vw.X), pk.iaddr(seg[i]), pk.iaddr(seg[i]),
                     addr(self.cross_mu, i * Cm), addr(self.cross_mu, i * Cm),
                     addr(cov, i * n_pad * n_pad), vw.N, T, vw.C, vw.C, vw.C, vw.C, n_pad, 1,
                     1.0 / (vw.N * T - 1), 0)
        d1, d2 = pk.add_descs(r1), pk.add_descs(r2)
        pk.upload()
        ctx.call('cpsd_colsum', pk.daddr(d1), nv, Cm)
        ctx.call(gram, pk.daddr(d2), nv, Cm, Cm)
        cd = ctypes_int_ptr(pk.iaddr(cdim))
        evals, evecs = self.eig_any(cov, n_pad, cd, 0, nv, 'cp')
        self.cross_k_dev = self.ws('cp_k', (nv,), I32)
        self._select_pca_k(evals, n_pad, cd, 0, self.cross_k_dev, 1, 0, nv, self.n_comp)
        self.cross_k = self.cross_k_dev.cpu().numpy().astype(np.int32)
        self.cross_evecs = evecs.clone()
        self.cross_evals = evals.clone()
        self.cross_npad = n_pad

    def _select_pca_k(self, evals, ld_e, n_dev, n_fixed, k_out, stride, offset, nprob, n_comp):
        if isinstance(n_comp, (float, np.floating)) and 0 < n_comp < 1:
            mode, thr = 0, float(n_comp)
        else:
            mode, thr = 3, float(int(n_comp))
        self.ctx.call('cpsd_select_k', ptr(evals), ld_e, _p(n_dev), n_fixed, thr, mode, 1, 1 << 30,
                      ptr(k_out, offset), stride, nprob)

    # ------------------------------------------------------------------ public API
    def _lanes(self, n):
        """Independent execution lanes: shallow copies of the engine that share the resident
        patient data and fold-invariant tables but own their stream, workspaces, index packs and
        view-statistics cache, so consecutive batches can be in flight at the same time."""
        extra = getattr(self, '_extra_lanes', None)
        if extra is None:
            extra = self._extra_lanes = []
        while 1 + len(extra) < n:
            import copy
            ln = copy.copy(self)
            ln._extra_lanes = None                    # no reference cycles: engines must die by
            ln._ws, ln._vs, ln._tc_stage, ln._marks = {}, None, None, []   # refcount
            ln._sched = {}          # block-Jacobi schedules are uploaded on the lane's own stream
            ln._xc = None
            ln._vb = None
            ln._jc = None
            ln._tkc_key = None
            ln.packA, ln.packB = HostPack(self.ctx), HostPack(self.ctx)
            ln.packM = [HostPack(self.ctx), HostPack(self.ctx)]
            ln._pack_i = 0
            if getattr(self, '_tcp', None) is not None:
                ln._tcp = dict(self._tcp, cap=0, nq=0)   # shares the X split + maps, own L^T buffers
            ln.stream = _lane_stream(self.ctx.device, self.lane + 1 + len(extra))
            ln.lane_idx = self.lane + 1 + len(extra)
            ln.ctx = LaneContext(self.ctx.base, ln.stream)
            extra.append(ln)
        lanes = [self] + extra
        return lanes[:n]

    def _bag_seeds(self, folds, bag_seeds):
        """Bagging decoders: the per-fold estimator seeds BaggingClassifier.fit draws
        (``random_state.randint(MAX_INT, size=n_estimators)``, sklearn/ensemble/_bagging.py); given
        by the caller (who replays the script's RNG stream) or drawn here from numpy's global RNG,
        one draw of n_estimators seeds per fold in fold order."""
        if not self.decoder.startswith('bag_'):
            return None
        if bag_seeds is None:
            bag_seeds = [np.random.randint(np.iinfo(np.int32).max, size=self.n_estimators)
                         for _ in folds]
        bag_seeds = np.asarray(bag_seeds, dtype=np.int64).reshape(len(folds), self.n_estimators)
        return bag_seeds

    def run(self, folds, return_details=False, bag_seeds=None, rep=None):
        """folds: list of (train_idx, test_idx) into the target's trials.  Returns a dict with
        ``y_pred`` (list of arrays, one per fold) and per-fold diagnostics."""
        out = {'y_pred': [], 'k2': [], 'h2d_bytes': 0, 'd2h_bytes': 0}
        details = []
        nb = max(1, -(-len(folds) // self.max_batch))
        size = -(-len(folds) // nb)          # balanced batches (a short tail batch costs as much
        seeds = self._bag_seeds(folds, bag_seeds)                           # as a full one)
        batches = []
        rep = self._check_rep(rep, len(folds))
        for s0 in range(0, len(folds), size):
            b = _Batch(folds[s0:s0 + size])
            b.bag_seeds = None if seeds is None else seeds[s0:s0 + size]
            b.rep = None if rep is None else rep[s0:s0 + size]
            batches.append(b)
        results = [None] * len(batches)
        self._ensure_ready()
        # every batch is a generator that yields right before each blocking read-back; the
        # lanes are advanced round-robin, so while one lane waits for its GPU results the
        # host packs and queues the other lane's batch on its own stream
        with torch.cuda.stream(self.stream):
            self._tc_proj_ready(1)                     # split X / encode maps before the lanes fork
            if self._runs_done or len(batches) > 1:
                self._tg_rotate()
        self._runs_done += 1
        nl = 1 if (self.profile or len(batches) == 1) else min(self.n_lanes, len(batches))
        lanes = self._lanes(nl)
        ready = torch.cuda.Event()
        ready.record(self.stream)
        for ln in lanes[1:]:
            ln.stream.wait_event(ready)            # fold-invariant tables live on lane 0's stream
        todo = [list(range(li, len(batches), nl)) for li in range(nl)]
        gens = [None] * nl
        cur = [None] * nl
        wait = [None] * nl          # event recorded when the lane last yielded
        live = sum(len(t) for t in todo)
        while live:
            progressed = False
            for li, ln in enumerate(lanes):
                if gens[li] is None:
                    if not todo[li]:
                        continue
                    cur[li] = todo[li].pop(0)
                    with torch.cuda.stream(ln.stream):
                        gens[li] = ln._batch_start(batches[cur[li]], return_details)
                    wait[li] = None
                    progressed = True
                    continue
                # resume a lane only when the work it queued before yielding has finished:
                # its read-back then returns at once and the host never sits in one lane's
                # wait while another lane has nothing queued
                if wait[li] is not None and nl > 1 and not wait[li].query():
                    continue
                progressed = True
                with torch.cuda.stream(ln.stream):
                    try:
                        if next(gens[li]) == 'host':
                            wait[li] = None      # nothing queued: resumable at once
                        else:
                            wait[li] = torch.cuda.Event()
                            wait[li].record(ln.stream)
                    except StopIteration as e:
                        results[cur[li]] = e.value
                        gens[li] = None
                        live -= 1
            if not progressed:
                time.sleep(_IDLE_SLEEP)
        for ln in lanes:
            ln.stream.synchronize()
        for res in results:
            out['y_pred'] += res['y_pred']
            out['k2'] += res['k2']
            out['h2d_bytes'] += res['h2d_bytes']
            out['d2h_bytes'] += res['d2h_bytes']
            if return_details:
                details.append(res['details'])
        if return_details:
            out['details'] = details
        return out

    def _check_rep(self, rep, n):
        """Replica index of every fold: ascending (the folds of one replica are contiguous)."""
        if rep is None:
            assert self.J == 1 or n == 0, 'an engine with replicas needs rep= for its folds'
            return None
        rep = np.asarray(rep, dtype=np.int64)
        assert rep.shape == (n,) and (n == 0 or (rep.min() >= 0 and rep.max() < self.J))
        assert (np.diff(rep) >= 0).all(), 'folds must be grouped by replica'
        return rep

    def run_gen(self, folds, return_details=False, bag_seeds=None, rep=None):
        """Generator form of run() on this engine's own stream only (no extra lanes): yields
        before every blocking read-back, returns the result dict.  The caller advances it with
        this engine's stream current (cv_align_decode_stream keeps several jobs in flight)."""
        out = {'y_pred': [], 'k2': [], 'h2d_bytes': 0, 'd2h_bytes': 0}
        details = []
        nb = max(1, -(-len(folds) // self.max_batch))
        size = -(-len(folds) // nb)
        if self.method == 'mcca' and self.cross_rank is None:
            yield 'sync'                     # constructor work still in flight
        self._ensure_ready()
        if self._runs_done or nb > 1:
            self._tg_rotate()
        self._runs_done += 1
        seeds = self._bag_seeds(folds, bag_seeds)
        rep = self._check_rep(rep, len(folds))
        for s0 in range(0, len(folds), size):
            batch = _Batch(folds[s0:s0 + size])
            batch.bag_seeds = None if seeds is None else seeds[s0:s0 + size]
            batch.rep = None if rep is None else rep[s0:s0 + size]
            res = yield from self._batch_start(batch, return_details)
            out['y_pred'] += res['y_pred']
            out['k2'] += res['k2']
            out['h2d_bytes'] += res['h2d_bytes']
            out['d2h_bytes'] += res['d2h_bytes']
            if return_details:
                details.append(res['details'])
        if return_details:
            out['details'] = details
        return out

    def _mcca_start(self, batch, want_details, align_only=False):
        """Generator of one MCCA batch (nothing runs until it is advanced)."""
        pk = self.packM[self._pack_i]
        self._pack_i ^= 1
        return self._batch_mcca_gen(batch, want_details, align_only, pk)

    def _batch_once(self, batch, want_details):
        if self.method in ('mcca', 'jointpca'):
            return self._mcca_start(batch, want_details)
        from .engine_cca import batch_cca_gen
        return batch_cca_gen(self, batch, want_details)

    def _batch_start(self, batch, want_details):
        """Generator of one batch for this engine's method: speculative first (see
        speculative_topk), again without speculation when the top-k round was not acceptable."""
        self._spec = (self.speculative_topk and self.decoder == 'linear'
                      and not getattr(self, '_spec_failed', False))
        self._spec_check = None
        try:
            res = yield from self._batch_once(batch, want_details)
        except _TopkRetry:
            # this workload needs more than one round: no more speculation on this lane
            self._spec = False
            self._spec_failed = True
            self.stats['topk_retries'] = self.stats.get('topk_retries', 0) + 1
            res = yield from self._batch_once(batch, want_details)
        return res

    def _batch_mcca(self, batch, want_details, align_only=False):
        self._ensure_ready()
        with torch.cuda.stream(self.stream):
            return _drain(self._mcca_start(batch, want_details, align_only))

    def align_mcca(self, train_idx=None):
        """MCCA fit only (AlignMCCA.fit): loadings, view means and generalised eigenvalues for
        the target's ``train_idx`` trials (default: all) pooled with the cross patients."""
        tr = np.arange(self.views[0].N) if train_idx is None else np.asarray(train_idx)
        return self._batch_mcca([(tr, np.zeros(0, dtype=np.int64))], False, align_only=True)

    # ------------------------------------------------------------------ shared host prep
    def _target_tables(self, pk, batch):
        """Per fold: target class-mean CSR (classes present among the train trials).  Generator
        (yields 'host' every few folds so the scheduler can service other lanes)."""
        tv = self.views[0]
        tabs = []
        for bi, (tr, te) in enumerate(batch):
            if bi and bi % 24 == 0:
                yield 'host'
            tr = np.asarray(tr, dtype=np.int64)
            cls = tv.cls[tr]
            present = np.unique(cls)
            order = np.argsort(cls, kind='stable')
            counts = np.bincount(cls, minlength=len(self.vocab))[present]
            mptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
            tabs.append(dict(tr=tr, te=np.asarray(te, dtype=np.int64), present=present,
                             o_ptr=pk.add_ints(mptr), o_mem=pk.add_ints(tr[order]),
                             o_tr=pk.add_ints(tr * self.T), o_te=pk.add_ints(np.asarray(te) * self.T)))
        return tabs

    def _class_means_target(self, pk, tabs, B, Kmax, xaddr=None):
        tv, T = self.views[0], self.T
        cmT = self.ws('cmT', (B, Kmax * T, tv.C))
        recs = np.zeros(B, dtype=_lib.CLASS_MEAN_DESC)
        ib = pk.iaddr(0)
        recs['X'] = addr(tv.X) if xaddr is None else xaddr
        if not isinstance(tabs, dict):       # list of per-fold tables (CCA / none batches)
            tabs = dict(o_ptr=[tb['o_ptr'] for tb in tabs], o_mem=[tb['o_mem'] for tb in tabs],
                        Kp=[len(tb['present']) for tb in tabs])
        recs['member_ptr'] = ib + 4 * np.asarray(tabs['o_ptr'], dtype=np.int64)
        recs['members'] = ib + 4 * np.asarray(tabs['o_mem'], dtype=np.int64)
        recs['out'] = addr(cmT) + 4 * Kmax * T * tv.C * np.arange(B, dtype=np.int64)
        recs['nslot'] = tabs['Kp']
        recs['TC'] = T * tv.C
        return cmT, recs

    # ------------------------------------------------------------------ pooled stage
    def _pooled_stage(self, pk, B, Zall, n_pad, F, n_pool, n_te, o_npool, o_nall, o_ypool,
                      ypool_ld, n_te_max, want_details):
        """Decoder: PCA(decoder_var) on the pooled matrix + OvR linear SVM + prediction.
        Zall (B, n_pad, F): rows [0,n_pool) pooled train, [n_pool, n_pool+n_te) test."""
        ctx = self.ctx
        n_all = [a + b for a, b in zip(n_pool, n_te)]
        mu = self.ws('pool_mu', (B, F))
        r1 = np.zeros(B, dtype=_lib.COLSUM_DESC)
        r2 = np.zeros(B, dtype=_lib.GRAM_NT_DESC)
        Kall = self.ws('pool_K', (B, n_pad, n_pad))
        fi = np.arange(B, dtype=np.int64)
        zb = addr(Zall) + 4 * n_pad * F * fi
        npool = np.asarray(n_pool, dtype=np.int64)
        r1['A'], r1['segA'], r1['out'] = zb, pk.iaddr(pk.o_zero), addr(mu) + 4 * F * fi
        r1['nseg'], r1['seg_len'], r1['p'], r1['lda'] = 1, npool, F, F
        r1['alpha'] = 1.0 / npool
        r2['A'], r2['B'], r2['out'] = zb, zb, addr(Kall) + 4 * n_pad * n_pad * fi
        r2['m'] = r2['n'] = np.asarray(n_all, dtype=np.int64)
        r2['k'], r2['lda'], r2['ldb'], r2['ldo'], r2['sym'], r2['alpha'] = F, F, F, n_pad, 1, 1.0
        pk.r2_host = r2
        return r1, r2, mu, Kall

    def gram_tc(self, recs_host, nprob, nmax, elems_per_prob, mu=None, ldmu=0):
        """Tensor-core pooled Gram (tcgen05): needs the descriptor records on the host.  With
        ``mu`` the rows are centred inside the hi/lo split."""
        ctx = self.ctx
        nbytes = int(ctx.lib.cpsd_gram_nt_tc_ws_bytes(nprob))
        split = self.ws('tc_split', (2 * nprob * elems_per_prob,))
        maps = self.ws('tc_maps', (nbytes + 64,), torch.uint8)
        if getattr(self, '_tc_stage', None) is None or self._tc_stage.numel() < nbytes:
            self._tc_stage = torch.empty((nbytes + 64,), dtype=torch.uint8).pin_memory()
        recs = np.ascontiguousarray(recs_host)
        a = (maps.data_ptr() + 63) & ~63
        mid = None
        if getattr(self, 'profile', False) and self._marks and self._marks[-1][0] == 'pool_gram':
            # stage timers: the hi/lo split (HBM-bound) and the MMA kernel (tensor-bound) separately
            mid = torch.cuda.Event(enable_timing=True)
            mid.record(torch.cuda.current_stream(ctx.device))      # creates the handle
            ctx.lib.cpsd_gram_nt_tc_probe(ctypes.c_void_p(mid.cuda_event))
            self._marks[-1] = ('pool_split', self._marks[-1][1])
        try:
            if mu is not None:
                ctx.call('cpsd_gram_nt_tc_centered', ctypes.c_void_p(recs.ctypes.data), nprob, nmax, nmax,
                         ptr(split), split.numel(), ctypes.c_void_p(a),
                         ctypes.c_void_p(self._tc_stage.data_ptr()), ptr(mu), ldmu)
            else:
                ctx.call('cpsd_gram_nt_tc', ctypes.c_void_p(recs.ctypes.data), nprob, nmax, nmax, ptr(split),
                         split.numel(), ctypes.c_void_p(a), ctypes.c_void_p(self._tc_stage.data_ptr()))
        finally:
            if mid is not None:
                ctx.lib.cpsd_gram_nt_tc_probe(ctypes.c_void_p(0))
                self._marks.append(('pool_gram', mid))

    def _pooled_stage_run(self, *args):
        """Blocking form of _pooled_stage_run_gen (CCA / none batches)."""
        return _drain(self._pooled_stage_run_gen(*args))

    def _pooled_stage_run_gen(self, pk, d1, d2, B, Zall, mu, Kall, n_pad, F, n_pool, n_te, o_npool,
                              o_nall, o_ypool, n_te_max, want_details):
        ctx = self.ctx
        npool_dev = ctypes_int_ptr(pk.iaddr(o_npool))
        nall_dev = ctypes_int_ptr(pk.iaddr(o_nall))
        ncls = len(self.classes)
        self.mark('pool_center')
        ctx.call('cpsd_colsum', pk.daddr(d1), B, F)
        nmax = max(a + b for a, b in zip(n_pool, n_te))
        tc_gram = self.use_tc and F % 4 == 0
        if not tc_gram:
            ctx.call('cpsd_center_rows', ptr(Zall), F, n_pad * F, ptr(mu), F, nall_dev, 0, nmax, F, B)
        self.mark('pool_gram')
        if tc_gram:       # centring fused into the tf32 split of the tensor-core Gram
            self.gram_tc(pk.r2_host, B, nmax, n_pad * F, mu=mu, ldmu=F)
        else:
            ctx.call('cpsd_gram_nt', pk.daddr(d2), B, nmax, nmax)
        self.mark('pool_eig')
        Kte = self.ws('pool_Kte', (B, n_te_max, n_pad))
        ctx.call('cpsd_copy_rows', ptr(Kall), n_pad, n_pad * n_pad, ptr(Kte), n_pad,
                 n_te_max * n_pad, npool_dev, 0, n_te_max, n_pad, n_pad, B)
        evals = self.ws('pool_ev', (B, n_pad))
        k2 = self.ws('pool_k2', (B,), I32)
        kcap = min(n_pad, F)
        if isinstance(self.decoder_var, (float, np.floating)) and 0 < self.decoder_var < 1:
            mode, thr = 0, float(self.decoder_var)
        else:
            mode, thr = 3, float(int(self.decoder_var))
        ldv, sV = n_pad, n_pad * n_pad
        sweeps = None
        V = None
        # every workspace used after the first read-back is fetched before it: the host packs
        # the next batch during that wait and may grow same-named workspaces
        St = self.ws('pool_St', (B, kcap, n_pad))
        Ste = self.ws('pool_Ste', (B, kcap, n_te_max))
        yhat = self.ws('yhat', (B, n_te_max), I32)
        m = self.topk_block
        # 'auto': the top-k iteration first (it converges in one round whenever the retained
        # components are a small, well separated part of the spectrum -- also on small pools: 2-patient
        # CCA 1980 -> 3160 folds/s), the full block-Jacobi solver when it is not accepted; an engine
        # whose SMALL pools (n_pad <= 512) needed more than two rounds (flat noisy spectra) goes
        # straight to the full solver from then on
        slow = getattr(self, '_topk_slow', None)
        if slow is None:
            slow = self._topk_slow = {}
        use_topk = (self.pool_solver != 'full' and n_pad > 128 and mode in (0, 3)
                    and (mode == 0 or int(thr) <= m - 8)
                    and (self.pool_solver == 'topk'
                         or (not slow.get('pool') and min(n_pool) - 1 >= m and F >= m)))
        perm_p = ptr(None)
        if use_topk:
            got = yield from self.eig_topk(Kall, n_pad, npool_dev, B, m, evals, 'pool', thr, mode,
                                           kcap, k2)
            if got is not None:
                V, ldv, sV = got
            # (only where the full solver is the cheaper alternative: small pools)
            if n_pad <= 512 and (got is None or self.stats['topk'].get('rounds', 1) > 2):
                slow['pool'] = True
            self.mark('pool_eigvecs')
        if V is None:
            V = self.ws('pool_V', (B, n_pad, n_pad))
            if n_pad <= 128:
                self.eig_small(Kall, npool_dev, 0, B, n_pad, evals, V, n_pad)
            else:
                perm = self.ws('pool_perm', (B, n_pad), I32)
                sweeps = self.eig_block(Kall, n_pad, npool_dev, 0, B, evals, perm, 'pool')
            self.mark('pool_eigvecs')
            ctx.call('cpsd_select_k', ptr(evals), n_pad, npool_dev, 0, thr, mode, 1, kcap, ptr(k2),
                     1, B)
            if n_pad > 128:
                # eigenvectors of the k2 retained components only (rotation-log replay); one
                # small read-back sizes the replay grid to the columns actually kept
                yield 'sync'
                self._k2_max = max(int(k2.cpu().numpy().max()), 1)
                k_launch = min(kcap, _ceil(self._k2_max, 64))
                self.eig_vecs('pool', n_pad, B, perm, k2, 0, k_launch, V)
            else:
                self._k2_max = kcap
        self.mark('pool_scores')
        ctx.call('cpsd_scores_train', ptr(V), ldv, sV, ptr(evals), perm_p, n_pad,
                 ptr(k2), npool_dev, 0, max(n_pool), ptr(St), n_pad, kcap * n_pad, kcap, B)
        ctx.call('cpsd_scores_test', ptr(Kte), n_pad, n_te_max * n_pad, ptr(V), ldv,
                 sV, ptr(evals), perm_p, n_pad, ptr(k2), npool_dev, 0, n_te_max,
                 ptr(Ste), n_te_max, kcap * n_te_max, kcap, B)
        return evals, k2, St, Ste, V, sweeps, kcap

    def _svm_stage(self, pk2, B, St, Ste, k2, kcap, n_pad, n_pool, n_te, o_ypool, ypool_ld,
                   o_nte, n_te_max, ypool=None, batch=None):
        ctx = self.ctx
        ncls = len(self.classes)
        if self.decoder.startswith('bag_'):
            # everything is sized after the decoder PCA (the bootstrap index streams depend on its
            # component count): see _decode_bagged
            state = dict(bag=True, ypool=ypool, n_pool=list(n_pool),
                         seeds=getattr(batch, 'bag_seeds', None))
            return state, self.ws('svc_info', (B, 1, 2), I32), np.zeros(0, dtype=_lib.SVM_DESC)
        if self.decoder != 'linear':
            # C-SVC: no descriptors; the largest class pair of the batch sizes the solver's
            # shared memory
            valid = np.arange(ypool.shape[1])[None, :] < np.asarray(n_pool)[:, None]
            ci = np.searchsorted(self.classes, ypool[valid])
            fo = np.broadcast_to(np.arange(B)[:, None], ypool.shape)[valid]
            cnt = np.bincount(fo * ncls + ci, minlength=B * ncls).reshape(B, ncls)
            top = np.sort(cnt, axis=1)
            m_max = int((top[:, -1] + top[:, -2]).max())
            npair = ncls * (ncls - 1) // 2
            state = dict(m_max=m_max, coef=self.ws('svc_coef', (B, ncls - 1, n_pad), torch.float64),
                         rho=self.ws('svc_rho', (B, npair), torch.float64),
                         gamma=self.ws('svc_gamma', (B,), torch.float64),
                         K=self.ws('svc_K', (B, n_pad, n_pad)), perm=self.ws('svc_perm', (B, n_pad), I32),
                         off=self.ws('svc_off', (B, ncls + 1), I32),
                         sqn=self.ws('svc_sqn', (B, n_pad), torch.float64))
            info = self.ws('svc_info', (B, npair, 2), I32)
            return state, info, np.zeros(0, dtype=_lib.SVM_DESC)
        W = self.ws('svm_W', (B, ncls, kcap + 1), torch.float64)
        info = self.ws('svm_info', (B, ncls, 4), I32)
        recs = np.zeros(B * ncls, dtype=_lib.SVM_DESC)
        fi = np.repeat(np.arange(B, dtype=np.int64), ncls)
        ti = np.arange(B * ncls, dtype=np.int64)
        recs['St'] = addr(St) + 4 * kcap * n_pad * fi
        recs['y'] = pk2.iaddr(o_ypool) + 4 * ypool_ld * fi
        recs['k_dev'] = addr(k2) + 4 * fi
        recs['w'] = addr(W) + 8 * (kcap + 1) * ti
        recs['info'] = addr(info) + 16 * ti
        recs['n'] = np.repeat(np.asarray(n_pool, dtype=np.int64), ncls)
        recs['lds'] = n_pad
        recs['cls'] = np.tile(self.classes.astype(np.int64), B)
        recs['C'], recs['tol_dcd'], recs['tol_newton'] = self.Csvm, self.tol_dcd, self.tol_newton
        recs['max_newton'], recs['dcd_epochs'] = self.max_newton, self.dcd_epochs
        return W, info, recs

    def _decode(self, pk, d_svm, W, info, B, St, Ste, k2, kcap, n_pad, o_ypool, ypool_ld, o_npool,
                o_nte, n_te_max):
        """Fits the decoder of every fold of the batch on its pooled scores and labels the
        held-out trials; returns the (B, n_te_max) label tensor."""
        ctx = self.ctx
        ncls = len(self.classes)
        yhat = self.ws('yhat', (B, n_te_max), I32)
        nte_dev = ctypes_int_ptr(pk.iaddr(o_nte))
        if self.decoder == 'linear':
            # shared memory of the solver is sized by the largest k2 of the batch, not by its cap
            ctx.call('cpsd_svm_fit_ovr_ex', pk.daddr(d_svm), B * ncls, min(kcap, self._k2_max), n_pad,
                     int(self.dcd_epochs))
            ctx.call('cpsd_svm_predict_ovr', ptr(Ste), n_te_max, kcap * n_te_max, ptr(W), kcap + 1,
                     ncls * (kcap + 1), ptr(k2), 0, nte_dev, n_te_max, ptr(self.classes_dev), ncls,
                     ptr(yhat), ptr(None), B)
            return yhat
        sv = W
        if sv.get('bag'):
            return self._decode_bagged(pk, sv, B, St, Ste, k2, kcap, n_pad, o_ypool, ypool_ld, o_npool,
                                       o_nte, n_te_max, yhat)
        if self._k2_max > 1024:
            raise NotImplementedError('C-SVC decoders support at most 1024 decoder-PCA components '
                                      '(this batch keeps %d)' % self._k2_max)
        kid = 1 if self.decoder == 'svc_rbf' else 0
        npool_dev = ctypes_int_ptr(pk.iaddr(o_npool))
        ypool_dev = ctypes_int_ptr(pk.iaddr(o_ypool))
        ctx.call('cpsd_svc_kernel_matrix', ptr(St), n_pad, kcap * n_pad, ptr(k2), 0, npool_dev, 0, n_pad,
                 ypool_dev, ypool_ld, ptr(self.classes_dev), ncls, kid, self.svc_gamma, ptr(sv['gamma']),
                 ptr(sv['perm']), ptr(sv['off']), ptr(sv['sqn']), ptr(sv['K']), n_pad, n_pad * n_pad, B)
        ctx.call('cpsd_svc_fit_ovo', ptr(sv['K']), n_pad, n_pad * n_pad, ptr(sv['perm']), ptr(sv['off']),
                 npool_dev, 0, ncls, self.Csvm, int(self.class_weight == 'balanced'),
                 self.svc_tol, self.svc_max_iter, ptr(sv['coef']), n_pad, ptr(sv['rho']), ptr(info),
                 sv['m_max'], B)
        ctx.call('cpsd_svc_predict_ovo', ptr(St), n_pad, kcap * n_pad, ptr(Ste), n_te_max,
                 kcap * n_te_max, ptr(k2), 0, npool_dev, 0, n_pad, nte_dev, n_te_max, ypool_dev, ypool_ld,
                 ptr(self.classes_dev), ncls, kid, ptr(sv['gamma']), ptr(sv['coef']), n_pad,
                 ptr(sv['rho']), ptr(yhat), ptr(None), min(kcap, 1024), B)
        return yhat

    def _check_decoder(self, info, B):
        """Linear decoder: counts the one-vs-rest problems whose Newton iteration hit max_newton
        (stats['svm_unconverged']) and warns like sklearn does for liblinear."""
        if self.decoder != 'linear':
            return
        bad = int((info.view(B, -1, 4)[:, :, 3] != 0).sum().item())
        self.stats['svm_unconverged'] = self.stats.get('svm_unconverged', 0) + bad
        if bad:
            import warnings
            from sklearn.exceptions import ConvergenceWarning
            warnings.warn('linear SVM: %d one-vs-rest problem(s) did not reach the optimum in %d '
                          'Newton steps' % (bad, self.max_newton), ConvergenceWarning)

    def _decode_bagged(self, pk, sv, B, St, Ste, k2, kcap, n_pad, o_ypool, ypool_ld, o_npool, o_nte,
                       n_te_max, yhat):
        """BaggingClassifier(estimator=SVC(kernel), n_estimators) per fold
        (scripts/aligned_decode_svm.py:262-265): the bootstrap sample of estimator e of fold f is
        sklearn's index stream for seed[f][e] (``_generate_bagging_indices``: the feature draw --
        all features, a permutation -- comes first, so the stream depends on the decoder-PCA
        component count); the resampled problems run as B * n_estimators folds of the C-SVC
        kernels, their labels are combined by majority vote."""
        from sklearn.utils.random import sample_without_replacement
        ctx = self.ctx
        ncls = len(self.classes)
        E = self.n_estimators
        k2h = k2.cpu().numpy()
        n_pool, ypool, seeds = sv['n_pool'], sv['ypool'], sv['seeds']
        idx = np.zeros((B * E, n_pad), dtype=np.int32)
        for f in range(B):
            kf, nf = int(k2h[f]), int(n_pool[f])
            for e in range(E):
                # sklearn/ensemble/_bagging.py::_generate_bagging_indices with bootstrap_features=
                # False, bootstrap=True, max_features = all, max_samples = all: the feature draw (a
                # permutation of all features, irrelevant to the kernels) consumes the stream first
                rs = np.random.RandomState(int(seeds[f][e]))
                sample_without_replacement(kf, kf, random_state=rs)
                idx[f * E + e, :nf] = rs.randint(0, nf, nf)
        fe = np.repeat(np.arange(B), E)
        yb = ypool[fe[:, None], idx]
        valid = np.arange(n_pad)[None, :] < np.asarray(n_pool)[fe][:, None]
        ci = np.searchsorted(self.classes, yb[valid])
        po = np.broadcast_to(np.arange(B * E)[:, None], yb.shape)[valid]
        cnt = np.bincount(po * ncls + ci, minlength=B * E * ncls).reshape(B * E, ncls)
        top = np.sort(cnt, axis=1)
        m_max = int((top[:, -1] + top[:, -2]).max())
        kb = _ceil(max(int(k2h.max()), 1), 4)
        if kb > 1024:
            raise NotImplementedError('C-SVC decoders support at most 1024 decoder-PCA components')
        BE = B * E
        idx_d = ctx.upload(idx, np.int32)
        St_b = self.ws('bag_St', (BE, kb, n_pad))
        Ste_b = self.ws('bag_Ste', (BE, kb, n_te_max))
        y_b = self.ws('bag_y', (BE, n_pad), I32)
        k_b, n_b, nte_b = (self.ws('bag_' + t, (BE,), I32) for t in ('k', 'n', 'nte'))
        yhat_b = self.ws('bag_yhat', (BE, n_te_max), I32)
        npair = ncls * (ncls - 1) // 2
        coef = self.ws('svc_coef', (BE, ncls - 1, n_pad), torch.float64)
        rho = self.ws('svc_rho', (BE, npair), torch.float64)
        gam = self.ws('svc_gamma', (BE,), torch.float64)
        K = self.ws('svc_K', (BE, n_pad, n_pad))
        perm = self.ws('svc_perm', (BE, n_pad), I32)
        off = self.ws('svc_off', (BE, ncls + 1), I32)
        sqn = self.ws('svc_sqn', (BE, n_pad), torch.float64)
        info = self.ws('bag_info', (BE, npair, 2), I32)
        npool_dev = ctypes_int_ptr(pk.iaddr(o_npool))
        nte_dev = ctypes_int_ptr(pk.iaddr(o_nte))
        ypool_dev = ctypes_int_ptr(pk.iaddr(o_ypool))
        kid = 1 if self.decoder.endswith('rbf') else 0
        ctx.call('cpsd_bag_gather', ptr(St), n_pad, kcap * n_pad, ptr(Ste), n_te_max, kcap * n_te_max,
                 ypool_dev, ypool_ld, ptr(idx_d), n_pad, ptr(k2), npool_dev, nte_dev, E, kb, ptr(St_b),
                 n_pad, ptr(Ste_b), n_te_max, ptr(y_b), ptr(k_b), ptr(n_b), ptr(nte_b), B)
        ctx.call('cpsd_svc_kernel_matrix', ptr(St_b), n_pad, kb * n_pad, ptr(k_b), 0, ptr(n_b), 0, n_pad,
                 ptr(y_b), n_pad, ptr(self.classes_dev), ncls, kid, self.svc_gamma, ptr(gam), ptr(perm),
                 ptr(off), ptr(sqn), ptr(K), n_pad, n_pad * n_pad, BE)
        ctx.call('cpsd_svc_fit_ovo', ptr(K), n_pad, n_pad * n_pad, ptr(perm), ptr(off), ptr(n_b), 0, ncls,
                 self.Csvm, int(self.class_weight == 'balanced'), self.svc_tol, self.svc_max_iter,
                 ptr(coef), n_pad, ptr(rho), ptr(info), m_max, BE)
        ctx.call('cpsd_svc_predict_ovo', ptr(St_b), n_pad, kb * n_pad, ptr(Ste_b), n_te_max,
                 kb * n_te_max, ptr(k_b), 0, ptr(n_b), 0, n_pad, ptr(nte_b), n_te_max, ptr(y_b), n_pad,
                 ptr(self.classes_dev), ncls, kid, ptr(gam), ptr(coef), n_pad, ptr(rho), ptr(yhat_b),
                 ptr(None), kb, BE)
        ctx.call('cpsd_bag_vote', ptr(yhat_b), E, ptr(self.classes_dev), ncls, nte_dev, n_te_max,
                 ptr(yhat), B)
        self._bag_keep = idx_d
        return yhat

    # ------------------------------------------------------------------ MCCA batch
    def _batch_mcca_gen(self, batch, want_details, align_only, pk):
        """Generator: runs the host packing + table upload, yields once, then launches."""
        ctx, T, P, Cm = self.ctx, self.T, self.P, self.Cmax
        B = len(batch)
        joint = self.method == 'jointpca'
        jvar = joint and self.joint_var
        Q = int(self._joint_qcap) if jvar else int(self.n_comp)
        use_rank = (0 < self.pca_var < 1) and not joint
        R = Q if use_rank else Cm      # pca_var == 1: no rank reduction (mvlearn _mcca_gevp)
        tv = self.views[0]
        J = self.J
        rep = getattr(batch, 'rep', None)
        rep = np.zeros(B, dtype=np.int64) if rep is None else np.asarray(rep, dtype=np.int64)
        assert J == 1 or not (joint or align_only), 'replicas: MCCA decode batches only'
        xaddr0 = np.array([addr(self.rviews[j][0].X) for j in range(J)], dtype=np.int64)[rep]   # target X per fold
        launches0 = ctx.launches()
        t_pack = time.perf_counter()
        pk.reset()
        pk.o_zero = pk.add_ints([0])
        # ---- index tables, built for all folds at once (2-D numpy ops over (fold, trial) /
        # (fold, class) instead of per-fold Python loops)
        V = len(self.vocab)
        n_tr_a = np.array([len(tr) for tr, _ in batch], dtype=np.int64)
        n_te_a = np.array([len(te) for _, te in batch], dtype=np.int64)
        ntr_max, nte_max = int(n_tr_a.max()), max(int(n_te_a.max()), 1)
        TR = np.full((B, ntr_max), -1, dtype=np.int64)
        TE = np.full((B, nte_max), -1, dtype=np.int64)
        for f, (tr, te) in enumerate(batch):
            TR[f, :n_tr_a[f]] = tr
            TE[f, :n_te_a[f]] = te
        mtr, mte = TR >= 0, TE >= 0
        fold_of = np.arange(B, dtype=np.int64)
        cls2d = np.where(mtr, tv.cls[np.where(mtr, TR, 0)], V)          # padding sorts last
        members = np.take_along_axis(TR, np.argsort(cls2d, axis=1, kind='stable'), axis=1)
        counts = np.bincount((fold_of[:, None] * (V + 1) + cls2d).ravel(),
                             minlength=B * (V + 1)).reshape(B, V + 1)[:, :V]
        present = counts > 0                                             # (fold, class)
        Kp = present.sum(axis=1)
        # class-mean CSR pointers of every fold: [0, cumulative counts of its present classes]
        p_start = np.concatenate([[0], np.cumsum(Kp + 1)])[:-1]
        mptr = np.zeros(int((Kp + 1).sum()), dtype=np.int32)
        mptr[np.arange(int(Kp.sum())) + 1 + np.repeat(fold_of, Kp)] = np.cumsum(counts, axis=1)[present]
        o_ptr = pk.add_ints(mptr) + p_start
        o_mem = pk.add_ints(np.where(members >= 0, members, 0)) + fold_of * ntr_max
        o_trv = pk.add_ints(np.where(mtr, TR, 0) * T) + fold_of * ntr_max
        o_tev = pk.add_ints(np.where(mte, TE, 0) * T) + fold_of * nte_max
        tabs = dict(o_ptr=o_ptr, o_mem=o_mem, Kp=Kp)
        self.stats['host_pack_ms'] = self.stats.get('host_pack_ms', 0.0) + \
            1e3 * (time.perf_counter() - t_pack)
        yield 'host'        # packing is pure host work: let the scheduler service other lanes
        t_pack = time.perf_counter()
        # shared class sets (alignment classes present in the fold's train trials AND in every
        # cross patient); folds with the same set share everything that depends on it
        cmask = np.ones(V, dtype=bool)
        if P > 1:
            cmask[:] = False
            cmask[sorted(set.intersection(*self.cross_classes))] = True
        shared2d = present & cmask[None, :]
        Ksa = shared2d.sum(axis=1).astype(np.int64)
        if Ksa.min() == 0:
            raise ValueError('no alignment class is shared by all patients in some fold')
        # folds with the same (replica, shared class set) share everything that depends on it
        uniq, first, inv = np.unique(np.column_stack([rep, shared2d.astype(np.int64)]), axis=0,
                                     return_index=True, return_inverse=True)
        inv = np.asarray(inv).ravel()
        rep_u = uniq[:, 0]
        shared_u = [np.nonzero(u)[0].astype(np.int64) for u in uniq[:, 1:]]
        keys_u = [(int(r), sh.tobytes()) for r, sh in zip(rep_u, shared_u)]
        shared = [shared_u[u] for u in inv]
        Kmax, Ks = int(Kp.max()), Ksa.tolist()
        KTmax = int(Ksa.max()) * T
        # segment tables: rows of the shared classes inside each view's class-mean array
        o_seg = np.zeros((B, P), dtype=np.int64)
        slot_in_fold = np.cumsum(present, axis=1) - 1                    # class -> row of the fold's means
        o_seg[:, 0] = pk.add_ints(slot_in_fold[shared2d] * T) + \
            np.concatenate([[0], np.cumsum(Ksa)])[:-1]
        if P > 1:
            seg_u = np.array([[pk.add_ints(self.cm_row[v][sh] * T) for v in range(1, P)]
                              for sh in shared_u], dtype=np.int64)
            o_seg[:, 1:] = seg_u[inv]
        o_segdst = pk.add_ints(np.arange(int(Ksa.max()), dtype=np.int32) * T)
        cdims = np.array([vw.C for vw in self.views], dtype=np.int64)
        o_cdim = pk.add_ints(np.tile(cdims, B))
        n_padC = 128 if Cm <= 128 else _ceil(Cm, 128)
        slot = np.zeros((B, P), dtype=np.int64)
        if joint:
            coff = np.concatenate([[0], np.cumsum(cdims)]).astype(np.int64)   # channel offsets
            nJ = int(coff[-1])
            nJ_pad = _ceil(nJ, 128)
            o_nj = pk.add_ints(np.full(B, nJ, dtype=np.int32))
        else:
            ranks = np.zeros((B, P), dtype=np.int32)
            ranks[:, 1:] = self.cross_rank[rep]
            o_rank = pk.add_ints(ranks)
            # slot of every (fold, view) problem; solve list = targets + cache misses
            n_padC = 128 if Cm <= 128 else _ceil(Cm, 128)
            vs = self._view_slots(B, n_padC, Cm)
            slot = np.zeros((B, P), dtype=np.int64)
            slot[:, 0] = fold_of
            for attempt in range(2):
                solve = [(f, 0, f) for f in range(B)]
                solve_u = []                 # shared-class-set index of every cross problem
                pending = {}
                if P > 1:
                    rows_u = np.zeros((len(shared_u), P - 1), dtype=np.int64)
                    for u, ku in enumerate(keys_u):      # ku = (replica, class set)
                        for v in range(1, P):
                            key = (v, ku)
                            sl = vs['keys'].get(key)
                            if sl is None:
                                sl = vs['next'] + len(pending)
                                pending[key] = sl
                                solve.append((int(first[u]), v, sl))
                                solve_u.append(u)
                            rows_u[u, v - 1] = sl
                    slot[:, 1:] = rows_u[inv]
                if vs['next'] + len(pending) <= vs['cap']:
                    break
                vs['keys'].clear()           # cache full: start over (everything misses once)
                vs['next'] = vs['res']
            o_slot = pk.add_ints(slot)
            o_cds = pk.add_ints([self.views[v].C for _, v, _ in solve])
            # cross problems: centred scatter from the per-class statistics (_cross_class_stats)
            xplan = None
            ccs = getattr(self, 'ccs', None)
            if ccs is not None and n_padC == 128 and len(solve) > B:
                lists, nrw = [], []
                for (f, v, sl), u in zip(solve[B:], solve_u):
                    rows = self.cm_row[v][shared_u[u]]
                    lists.append(ccs['off'][v] + int(rep_u[u]) * ccs['ncl'][v] + rows)
                    nrw.append(len(rows) * T)
                xplan = dict(o_lptr=pk.add_ints(np.concatenate([[0], np.cumsum([len(l) for l in lists])])),
                             o_list=pk.add_ints(np.concatenate(lists)), o_nrows=pk.add_ints(nrw))
            # warm-start plan of the view solves: the first scatter matrix solved for a (replica,
            # patient) pair is solved cold and its eigenvectors become the pair's basis; every
            # other problem of the pair is rotated into that basis and starts from it
            wplan = None
            if self.warm_start and n_padC == 128:
                vb = self._view_bases()
                s_f = np.array([t[0] for t in solve], dtype=np.int64)
                code = rep[s_f] * P + np.array([t[1] for t in solve], dtype=np.int64)
                cold, warm_i, base_of, tab = plan_view_solves(code, vb['tab'], vb['n'])
                s_slot = np.array([t[2] for t in solve], dtype=np.int32)
                wplan = dict(ncold=len(cold), nwarm=len(warm_i), tab=tab,
                             o_selc=pk.add_ints(cold), o_cslot=pk.add_ints(s_slot[cold]),
                             o_selw=pk.add_ints(warm_i), o_basew=pk.add_ints(base_of[warm_i]),
                             o_v0=pk.add_ints(base_of), o_out=pk.add_ints(s_slot))
            # cross-block cache slot of every fold (one slot per shared class set)
            XR = (P - 1) * R
            xc = self._xcache(R, XR)
            newx = []                       # (cache slot, first fold with that class set)
            xs_of = {}
            if len(xc['keys']) + sum(ku not in xc['keys'] for ku in keys_u) > xc['cap']:
                xc['keys'].clear()           # cache full: start over
            xs_u = np.zeros(len(shared_u), dtype=np.int32)
            for u, ku in enumerate(keys_u):
                xs = xc['keys'].get(ku)
                if xs is None:
                    xs = len(xc['keys']) + len(xs_of)
                    xs_of[ku] = xs
                    newx.append((xs, int(first[u])))
                xs_u[u] = xs
            xslot = xs_u[inv]
            keys = [keys_u[u] for u in inv]
            o_xslot = pk.add_ints(xslot)
        self.stats['host_pack_ms'] = self.stats.get('host_pack_ms', 0.0) + \
            1e3 * (time.perf_counter() - t_pack)
        yield 'host'
        t_pack = time.perf_counter()
        downdate = (not joint) and use_rank and getattr(self, 'tg', None) is not None and n_padC == 128
        if downdate:
            # 'all trials minus held-out trials' equals the train-set Gram only when train and
            # test partition the target's trials exactly (plain K-fold units); subsampled train
            # sets, fit-only calls and inner folds of a nested search sum the train list instead
            flat = np.concatenate([np.nonzero(mtr)[0] * tv.N + TR[mtr], np.nonzero(mte)[0] * tv.N + TE[mte]])
            hits = np.bincount(flat, minlength=B * tv.N).reshape(B, tv.N)
            use_te = bool((hits == 1).all()) and int(n_te_a.sum()) <= int(n_tr_a.sum())
            # rows of self.tg['trial']: replica * N + trial, then J * N + replica = minus the sum of
            # the replica's trials; a fold's Gram is -(listed rows) (held-out form) or +(train rows)
            La, Lm = (TE, mte) if use_te else (TR, mtr)
            Lf = np.where(Lm, La + rep[:, None] * tv.N, -1)
            if use_te:
                Lf = np.concatenate([(J * tv.N + rep)[:, None], Lf], axis=1)
            Lk = Lf >= 0
            o_lptr = pk.add_ints(np.concatenate([[0], np.cumsum(Lk.sum(axis=1))]))
            o_list = pk.add_ints(Lf[Lk])
        # pooled layout
        n_tr, n_te = n_tr_a.tolist(), n_te_a.tolist()
        cross_N = [self.views[v].N for v in range(1, P)]
        n_pool = [(nt if self.tar_in_train else 0) + sum(cross_N) for nt in n_tr]
        n_te_max = max(n_te)
        n_pad = _ceil(max(a + b for a, b in zip(n_pool, n_te)), 128)
        F = T * Q
        tc_proj = (not align_only) and self._tc_proj_ready(Q)
        assert J == 1 or tc_proj, 'replicas need the tensor-core projection (use_tensor_cores=True)'
        ycross = getattr(self, '_ycross', None)
        if ycross is None:
            ycross = self._ycross = (np.concatenate([self.views[v].y for v in range(1, P)])
                                     if P > 1 else np.zeros(0, dtype=np.int64)).astype(np.int32)
        ypool = np.zeros((B, n_pad), dtype=np.int32)
        row0 = n_tr_a.copy() if self.tar_in_train else np.zeros(B, dtype=np.int64)
        fr, fc = np.nonzero(mtr)                    # (fold, rank inside the fold's train list)
        if self.tar_in_train:
            ypool[fr, fc] = tv.y[TR[mtr]]
        if len(ycross):
            ypool[fold_of[:, None], row0[:, None] + np.arange(len(ycross))[None, :]] = ycross[None, :]
        if not tc_proj:
            o_pooldst = np.zeros((B, P), dtype=np.int64)
            o_allseg = [pk.add_ints(np.arange(self.views[v].N, dtype=np.int32) * T)
                        for v in range(P)]
            o_tedst = []
            for f in range(B):
                row = 0
                if self.tar_in_train:
                    o_pooldst[f, 0] = pk.add_ints((row + np.arange(n_tr[f])) * T)
                    row += n_tr[f]
                for v in range(1, P):
                    o_pooldst[f, v] = pk.add_ints((row + np.arange(self.views[v].N)) * T)
                    row += self.views[v].N
                o_tedst.append(pk.add_ints((row + np.arange(n_te[f])) * T))
        else:
            # destination trial row of every (fold, view, trial) in the fold's pooled matrix
            Nmax = max(vw.N for vw in self.views)
            cbase = -np.ones((P, Nmax), dtype=np.int64)
            off = 0
            for v in range(1, P):
                cbase[v, :self.views[v].N] = off + np.arange(self.views[v].N)
                off += self.views[v].N
            dst = np.where(cbase[None] >= 0, cbase[None] + row0[:, None, None], -1).astype(np.int32)
            if self.tar_in_train:
                dst[fr, 0, TR[mtr]] = fc
            er, ec = np.nonzero(mte)
            dst[er, 0, TE[mte]] = np.asarray(n_pool, dtype=np.int64)[er] + ec
            o_dst = pk.add_ints(dst)
        o_ypool = pk.add_ints(ypool)
        o_npool = pk.add_ints(n_pool)
        o_nall = pk.add_ints([a + b for a, b in zip(n_pool, n_te)])
        o_nte = pk.add_ints(n_te)
        self.stats['host_pack_ms'] = self.stats.get('host_pack_ms', 0.0) + \
            1e3 * (time.perf_counter() - t_pack)
        yield 'host'
        t_pack = time.perf_counter()
        pk.reserve_ints()
        ib = pk.iaddr(0)

        # ---- descriptors, stage A (filled column-wise: one numpy op per field)
        cmT, r_cm = self._class_means_target(pk, tabs, B, Kmax, xaddr=xaddr0)
        fi = np.arange(B, dtype=np.int64)
        # base address of the class-mean array of every (fold, view)
        cmb = np.empty((B, P), dtype=np.int64)
        cmb[:, 0] = addr(cmT) + 4 * Kmax * T * tv.C * fi
        for v in range(1, P):
            cmb[:, v] = addr(self.cm[v]) + 4 * self.cm[v][0].numel() * rep
        segb = ib + 4 * o_seg                       # (B, P) segment-table addresses
        Ksa = np.asarray(Ks, dtype=np.int64)
        o_tr = ib + 4 * o_trv
        if joint:
            # ---- JointPCA (alignment/JointPCA.py:165-211): per fold the Gram of the channel-
            # concatenated class averages (fp64, upper blocks), its column sums, the PCA of the
            # concatenation and one least-squares read-in matrix per patient
            mub = np.zeros((B, P), dtype=np.int64)             # no centring at transform time
            Gj = self.ws('j_G', (B, nJ, nJ), torch.float64)
            sj = self.ws('j_s', (B, nJ))
            # The blocks between cross patients (and their column sums) depend only on the
            # shared class set, not on the fold: they are computed once per set into a cache
            # slot and copied into every fold's Gram; per fold only the target's row of blocks
            # is new (the reference refits everything for every fold).
            jc = getattr(self, '_jc', None)
            if jc is None or jc['nJ'] != nJ or jc['cap'] < len(keys_u):
                cap = max(8, len(keys_u) + 8)
                jc = self._jc = dict(nJ=nJ, cap=cap, keys={}, G=self.ctx.zeros((cap, nJ, nJ), torch.float64),
                                     s=self.ctx.zeros((cap, nJ)))
            if len(jc['keys']) + len(set(keys_u) - set(jc['keys'])) > jc['cap']:
                jc['keys'].clear()
            miss_u = [u for u, ku in enumerate(keys_u) if ku not in jc['keys']]
            for u in miss_u:
                jc['keys'][keys_u[u]] = len(jc['keys'])
            jslot = np.array([jc['keys'][ku] for ku in keys_u], dtype=np.int64)[inv]     # per fold
            first = np.asarray(first).ravel()
            tp = [(0, v) for v in range(P)]                                  # target row, every fold
            xp = [(u, v) for u in range(1, P) for v in range(u, P)]          # cross blocks, misses only
            ff = np.concatenate([np.repeat(fi, len(tp)), np.repeat(first[miss_u], len(xp))]).astype(np.int64)
            uu = np.concatenate([np.tile([a for a, _ in tp], B), np.tile([a for a, _ in xp], len(miss_u))]).astype(np.int64)
            vv = np.concatenate([np.tile([b for _, b in tp], B), np.tile([b for _, b in xp], len(miss_u))]).astype(np.int64)
            n_t = B * len(tp)
            r_jg = np.zeros(len(ff), dtype=_lib.GRAM_TN_DESC)
            r_jg['A'], r_jg['B'] = cmb[ff, uu], cmb[ff, vv]
            r_jg['segA'], r_jg['segB'] = segb[ff, uu], segb[ff, vv]
            out = addr(Gj) + 8 * (nJ * nJ * ff + coff[uu] * nJ + coff[vv])
            out[n_t:] = addr(jc['G']) + 8 * (nJ * nJ * jslot[ff[n_t:]] + coff[uu[n_t:]] * nJ + coff[vv[n_t:]])
            r_jg['out'] = out
            r_jg['nseg'], r_jg['seg_len'] = Ksa[ff], T
            r_jg['p'] = r_jg['lda'] = cdims[uu]
            r_jg['q'] = r_jg['ldb'] = cdims[vv]
            r_jg['ldo'], r_jg['sym'], r_jg['alpha'] = nJ, (uu == vv).astype(np.int32), 1.0
            npair_launch = len(ff)
            # column sums: the target's per fold, the cross patients' per missing set
            sf = np.concatenate([fi, np.repeat(first[miss_u], P - 1)]).astype(np.int64)
            sv = np.concatenate([np.zeros(B, dtype=np.int64), np.tile(np.arange(1, P, dtype=np.int64), len(miss_u))])
            r_js = np.zeros(len(sf), dtype=_lib.COLSUM_DESC)
            r_js['A'], r_js['segA'] = cmb[sf, sv], segb[sf, sv]
            so = addr(sj) + 4 * (nJ * sf + coff[sv])
            so[B:] = addr(jc['s']) + 4 * (nJ * jslot[sf[B:]] + coff[sv[B:]])
            r_js['out'] = so
            r_js['nseg'], r_js['seg_len'], r_js['p'], r_js['lda'] = Ksa[sf], T, cdims[sv], cdims[sv]
            r_js['alpha'] = 1.0
            ncs_launch = len(sf)
            jslot_dev = torch.from_numpy(jslot).pin_memory().to(ctx.device, non_blocking=True)
            d_cm = pk.add_descs(r_cm)
            d_jg, d_js = pk.add_descs(r_jg), pk.add_descs(r_js)
        else:
            Gt, gram_c = self.scatter('m_Gt', B, n_padC)
            nS = len(solve)
            mu = vs['mu']
            cov, _ = self.scatter('m_cov', nS, n_padC)
            esz = cov.element_size()
            mub = addr(mu) + 4 * Cm * slot               # (B, P) mean vectors (slots)
            r_gt = np.zeros(B, dtype=_lib.GRAM_TN_DESC)
            r_gt['A'] = r_gt['B'] = xaddr0
            r_gt['segA'] = r_gt['segB'] = o_tr
            r_gt['out'] = addr(Gt) + Gt.element_size() * n_padC * n_padC * fi
            r_gt['nseg'], r_gt['seg_len'] = n_tr, T
            r_gt['p'] = r_gt['q'] = r_gt['lda'] = r_gt['ldb'] = tv.C
            r_gt['ldo'], r_gt['sym'], r_gt['alpha'] = n_padC, 1, 1.0
            sf = np.array([t[0] for t in solve], dtype=np.int64)
            sv = np.array([t[1] for t in solve], dtype=np.int64)
            ssl = np.array([t[2] for t in solve], dtype=np.int64)
            sC = cdims[sv]
            r_mu = np.zeros(nS, dtype=_lib.COLSUM_DESC)
            r_mu['A'], r_mu['segA'], r_mu['out'] = cmb[sf, sv], segb[sf, sv], addr(mu) + 4 * Cm * ssl
            r_mu['nseg'], r_mu['seg_len'], r_mu['p'], r_mu['lda'] = Ksa[sf], T, sC, sC
            r_mu['alpha'] = 1.0 / (Ksa[sf] * T)
            r_cov = np.zeros(nS, dtype=_lib.GRAM_TN_DESC)
            r_cov['A'] = r_cov['B'] = cmb[sf, sv]
            r_cov['segA'] = r_cov['segB'] = segb[sf, sv]
            r_cov['muA'] = r_cov['muB'] = addr(mu) + 4 * Cm * ssl
            r_cov['out'] = addr(cov) + esz * n_padC * n_padC * np.arange(nS, dtype=np.int64)
            r_cov['nseg'], r_cov['seg_len'] = Ksa[sf], T
            r_cov['p'] = r_cov['q'] = r_cov['lda'] = r_cov['ldb'] = sC
            r_cov['ldo'], r_cov['sym'], r_cov['alpha'] = n_padC, 1, 1.0
            d_cm, d_gt = pk.add_descs(r_cm), pk.add_descs(r_gt)
            d_mu, d_cov = pk.add_descs(r_mu), pk.add_descs(r_cov)
            # reduced views
            Vr = self.ws('m_Vr', (B * P, Cm, R))
            d2 = self.ws('m_d2', (B * P, R))
            r_eff = self.ws('m_reff', (B * P,), I32)
            self.stats['host_pack_ms'] = self.stats.get('host_pack_ms', 0.0) + \
                1e3 * (time.perf_counter() - t_pack)
            yield 'host'
            t_pack = time.perf_counter()
            # reduced coordinates of the condition averages: the target's per fold (Zt), the cross
            # patients' once per shared class set (Zx, cached with their cross-scatter block Gxx)
            nX = len(newx)
            KTc = xc['KT']
            Zt = self.ws('m_Zt', (B, KTmax, R))
            Gtx = self.ws('m_Gtx', (B, R, P * R))
            Zx, Gxx = xc['Zx'], xc['Gxx']
            npz = B + nX * (P - 1)
            r_pz = np.zeros(npz, dtype=_lib.PROJ_DESC)
            # targets
            r_pz['X'][:B], r_pz['seg_src'][:B] = cmb[:, 0], segb[:, 0]
            r_pz['mu'][:B], r_pz['W'][:B] = mub[:, 0], addr(Vr) + 4 * Cm * R * P * fi
            r_pz['Y'][:B] = addr(Zt) + 4 * KTmax * R * fi
            r_pz['nseg'][:B], r_pz['C'][:B], r_pz['ldx'][:B], r_pz['ldy'][:B] = Ksa, cdims[0], cdims[0], R
            j = B
            for xs, f in newx:
                for v in range(1, P):
                    r_pz[j] = (cmb[f, v], segb[f, v], 0, mub[f, v], addr(Vr, (f * P + v) * Cm * R),
                               addr(Zx, xs * KTc * XR + (v - 1) * R), Ks[f], T, cdims[v], R, cdims[v],
                               R, XR, 0)
                    j += 1
            r_pz['seg_dst'], r_pz['seg_len'], r_pz['q'], r_pz['ldw'] = ib + 4 * o_segdst, T, R, R
            ng = (2 * B if XR else B) + nX
            r_g = np.zeros(ng, dtype=_lib.GRAM_TN_DESC)
            zt = addr(Zt) + 4 * KTmax * R * fi
            r_g['A'][:B] = r_g['B'][:B] = zt
            r_g['out'][:B] = addr(Gtx) + 4 * R * P * R * fi
            r_g['seg_len'][:B] = Ksa * T
            r_g['p'][:B] = r_g['q'][:B] = r_g['lda'][:B] = r_g['ldb'][:B] = R
            r_g['sym'][:B] = 1
            if XR:
                r_g['A'][B:2 * B] = zt
                r_g['B'][B:2 * B] = addr(Zx) + 4 * KTc * XR * xslot.astype(np.int64)
                r_g['out'][B:2 * B] = addr(Gtx) + 4 * (R * P * R * fi + R)
                r_g['seg_len'][B:2 * B] = Ksa * T
                r_g['p'][B:2 * B] = r_g['lda'][B:2 * B] = R
                r_g['q'][B:2 * B] = r_g['ldb'][B:2 * B] = XR
                j = 2 * B
                for xs, f in newx:
                    zx = addr(Zx, xs * KTc * XR)
                    r_g[j] = (zx, zx, 0, 0, 0, 0, addr(Gxx, xs * XR * XR), 1, Ks[f] * T, XR, XR, XR, XR,
                              XR, 1, 1.0, 0)
                    j += 1
            r_g['segA'] = r_g['segB'] = pk.iaddr(pk.o_zero)
            r_g['nseg'], r_g['alpha'] = 1, 1.0
            r_g['ldo'][:2 * B if XR else B] = P * R
            d_pz, d_g = pk.add_descs(r_pz), pk.add_descs(r_g)
            # pooled projection
            # size of the reduced GEVP: the target's rank is only known on the device (<= R), the
            # cross patients' ranks are fold-invariant and known here
            if use_rank:
                n_m_max = R + int(np.minimum(self.cross_rank, R).sum(axis=1).max())
            else:
                n_m_max = sum(min(R, vw.C) for vw in self.views)
            n_padM = 128 if n_m_max <= 128 else _ceil(n_m_max, 128)
        L = self.ws('m_L', (B * P, Cm, Q))
        if not align_only:
            Zall = self.ws('pool_Z', (B, n_pad, F))
            if not tc_proj:
                r_pp = np.zeros(B * P + B, dtype=_lib.PROJ_DESC)
                for f in range(B):
                    for v in range(P):
                        vw = self.views[v]
                        i = f * P + v
                        if v == 0:
                            nseg = n_tr[f] if self.tar_in_train else 0
                            src = pk.iaddr(int(o_trv[f]))
                        else:
                            nseg, src = vw.N, pk.iaddr(o_allseg[v])
                        r_pp[i] = (addr(vw.X), src, pk.iaddr(o_pooldst[f, v]), int(mub[f, v]),
                                   addr(L, i * Cm * Q), addr(Zall, f * n_pad * F), nseg, T, vw.C, Q,
                                   vw.C, Q, Q, 0)
                    r_pp[B * P + f] = (addr(tv.X), pk.iaddr(int(o_tev[f])), pk.iaddr(o_tedst[f]),
                                       int(mub[f, 0]), addr(L, f * P * Cm * Q),
                                       addr(Zall, f * n_pad * F), n_te[f], T, tv.C, Q, tv.C, Q, Q, 0)
                d_pp = pk.add_descs(r_pp)
            r1, r2, pmu, Kall = self._pooled_stage(pk, B, Zall, n_pad, F, n_pool, n_te, o_npool,
                                                   o_nall, o_ypool, n_pad, n_te_max, want_details)
            d_p1, d_p2 = pk.add_descs(r1), pk.add_descs(r2)
            kcap = min(n_pad, F)
            # SVM descriptors need k2 / St addresses only (known now)
            St = self.ws('pool_St', (B, kcap, n_pad))
            k2 = self.ws('pool_k2', (B,), I32)
            W, info, r_svm = self._svm_stage(pk, B, St, None, k2, kcap, n_pad, n_pool, n_te, o_ypool,
                                             n_pad, o_nte, n_te_max, ypool, batch=batch)
            d_svm = pk.add_descs(r_svm)
        pk.upload()
        self.stats['host_pack_ms'] = self.stats.get('host_pack_ms', 0.0) + \
            1e3 * (time.perf_counter() - t_pack)
        yield 'host'
        t_pack = time.perf_counter()

        # ---- launches
        cdim_dev = ctypes_int_ptr(pk.iaddr(o_cdim))
        rank_dev = self.ws('m_rank', (B * P,), I32)
        # class means of the target's train trials
        self.mark('class_mean')
        ctx.call('cpsd_class_mean', pk.daddr(d_cm), B, Kmax, T * tv.C)
        status = self.ws('m_status', (B,), I32)
        if joint:
            # ---- JointPCA fit of every fold
            self.mark('align_scatter_eig')
            ctx.call('cpsd_colsum', pk.daddr(d_js), ncs_launch, Cm)
            ctx.call('cpsd_gram_tn_f64', pk.daddr(d_jg), npair_launch, Cm, Cm)
            if P > 1:       # cross blocks / sums of every fold from its set's cache slot (device copy)
                c1 = int(coff[1])
                Gj[:B, c1:, c1:] = jc['G'][jslot_dev][:, c1:, c1:]
                sj[:B, c1:] = jc['s'][jslot_dev][:, c1:]
            nrows_dev = self.ws('j_nrows', (B,), I32)
            nrows_dev.copy_(torch.from_numpy((Ksa * T).astype(np.int32)).pin_memory(), non_blocking=True)
            covj = self.ws('j_cov', (B, nJ_pad, nJ_pad))
            ctx.call('cpsd_joint_cov', ptr(Gj), ptr(sj), ptr(nrows_dev), nJ, ptr(covj), nJ_pad, B)
            self.mark('mcca_gevp')
            nj_dev = ctypes_int_ptr(pk.iaddr(o_nj))
            evj = self.ws('j_ev', (B, nJ_pad))
            kj = self.ws('j_k', (B,), I32)
            Vj = None
            mj = self.topk_block
            if nJ_pad > 128 and Q <= mj - 8 and self.pool_solver != 'full':
                # above 1000 concatenated channels sklearn's PCA itself switches to randomized SVD
                # (7 power iterations, random start: decomposition/_pca.py 'auto' policy), so
                # the trailing noise-floor components are not defined to better than that
                jtol = self.topk_tol if nJ <= 1000 else max(self.topk_tol, 1e-3)
                jthr, jmode = (float(self.n_comp), 0) if jvar else (float(Q), 3)
                got = yield from self.eig_topk(covj, nJ_pad, nj_dev, B, mj, evj, 'joint', jthr, jmode,
                                               nJ_pad, kj, tol=jtol,
                                               gap_tol=None if nJ <= 1000 else float('inf'))
                if got is not None:
                    Vj, ldvj, sVj = got
            if Vj is None:
                evj, Vj = self.eig_any(covj, nJ_pad, nj_dev, 0, B, 'jointf', ncols=_ceil(Q, 64))
                ldvj, sVj = nJ_pad, nJ_pad * nJ_pad
                if jvar:
                    ctx.call('cpsd_select_k', ptr(evj), nJ_pad, nj_dev, 0, float(self.n_comp), 0, 1,
                             1 << 30, ptr(kj), 1, B)
            rhs = self.ws('j_rhs', (B * P, Cm, Q), torch.float64)
            coff_dev = self.ws('j_coff', (P + 1,), I32)
            coff_dev.copy_(torch.from_numpy(coff.astype(np.int32)).pin_memory(), non_blocking=True)
            ctx.call('cpsd_joint_rhs', ptr(Gj), ptr(sj), ptr(nrows_dev), nJ, ptr(Vj), ldvj, sVj,
                     ptr(coff_dev), P, Cm, Q, ptr(rhs), Q, Cm * Q, B)
            status.zero_()
            L.zero_()
            stj = self.ws('j_status', (B * P,), I32)
            stj.zero_()
            for v in range(P):           # S_p W_p = rhs_p with S_p = G's diagonal block of patient p
                C = int(cdims[v])
                if C > 128:          # factor in an L2-resident workspace instead of shared memory
                    cw = self.ws('j_cholws', (B * C * (C + 1),), torch.float64)
                    ctx.call('cpsd_chol_solve_f64_ws', ptr(Gj, int(coff[v]) * nJ + int(coff[v])), nJ,
                             nJ * nJ, C, ptr(rhs, v * Cm * Q), Q, P * Cm * Q, Q, ptr(L, v * Cm * Q), Q,
                             P * Cm * Q, ptr(stj, v), ptr(cw), B)
                    continue
                ctx.call('cpsd_chol_solve_f64', ptr(Gj, int(coff[v]) * nJ + int(coff[v])), nJ, nJ * nJ, C,
                         ptr(rhs, v * Cm * Q), Q, P * Cm * Q, Q, ptr(L, v * Cm * Q), Q, P * Cm * Q,
                         ptr(stj, v), B)
            if jvar:       # every fold keeps its own component count
                ctx.call('cpsd_mask_cols', ptr(L), Q, Cm * Q, Cm, Q, ptr(kj), P, B * P)
            if align_only:
                torch.cuda.synchronize(self.ctx.device)
                return dict(loadings=L.view(B, P, Cm, Q).cpu().numpy(), shared=[s_.copy() for s_ in shared],
                            evals=evj[:, :Q].cpu().numpy(), status=stj.cpu().numpy(),
                            k=kj.cpu().numpy() if jvar else np.full(B, Q))
        else:
            # signal ranks (cross ranks are fold-invariant and come with the int table)
            ctx.call('cpsd_copy_rows', ctypes_int_ptr(pk.iaddr(o_rank)), B * P, 0, ptr(rank_dev),
                     B * P, 0, ptr(None), 0, 1, B * P, 1, 1)
            self.mark('align_scatter_eig')
            if Cm < n_padC:
                cov.zero_()
                Gt.zero_()
            rank_done = None
            if use_rank and downdate and self.side_streams and not self.profile and B <= 64:
                # small batches (one job per call: 20 problems per launch) leave most SMs idle, and
                # the signal-rank solve is independent of the view solves until cpsd_mcca_mask_idx:
                # it runs on the lane's side stream (same workspaces, explicit stream handle only).
                # Full batches fill the SMs with one launch: nothing to overlap (measured)
                ev_t = self.ws('mrk_ev', (B, n_padC))
                sd = _lane_stream(ctx.device, 16 + self.lane_idx)
                sx = LaneContext(ctx.base, sd)
                fork = torch.cuda.Event()
                fork.record(ctx.torch_stream)
                sd.wait_event(fork)
                sx.call('cpsd_sum_mats_f64', ptr(None), ptr(self.tg['trial']), 128 * 128,
                        ctypes_int_ptr(pk.iaddr(o_lptr)), ctypes_int_ptr(pk.iaddr(o_list)),
                        -1.0 if use_te else 1.0, ptr(Gt), 128 * 128, 128 * 128, B)
                sx.call('cpsd_eig_sym_small_f64', ptr(Gt), n_padC, n_padC * n_padC, ptr(None), tv.C, B,
                        ptr(ev_t), n_padC, ptr(None), n_padC, n_padC * n_padC, self.eig_sweeps + 6, 1e-10,
                        ptr(None))
                sx.call('cpsd_select_k', ptr(ev_t), n_padC, ptr(None), tv.C, float(self.pca_var), 1,
                        0, 1 << 30, ptr(rank_dev), P, B)
                rank_done = torch.cuda.Event()
                rank_done.record(sd)
            elif use_rank:
                if downdate:     # train-set Gram = all-trials Gram - held-out trials (or sum of train)
                    ctx.call('cpsd_sum_mats_f64', ptr(None),
                             ptr(self.tg['trial']), 128 * 128, ctypes_int_ptr(pk.iaddr(o_lptr)),
                             ctypes_int_ptr(pk.iaddr(o_list)), -1.0 if use_te else 1.0, ptr(Gt),
                             128 * 128, 128 * 128, B)
                else:
                    self.gram_scatter(gram_c, pk.daddr(d_gt), B, tv.C, tv.C, Gt, 3)
                ev_t, _ = self.eig_any(Gt, n_padC, ptr(None), tv.C, B, 'mrk', vecs=False)
                ctx.call('cpsd_select_k', ptr(ev_t), n_padC, ptr(None), tv.C, float(self.pca_var), 1,
                         0, 1 << 30, ptr(rank_dev), P, B)
            # per-view centred scatter of the condition averages + eigen-decomposition, for the
            # target of every fold and for the cross-patient problems not solved before
            nM, s0 = nS - B, vs['next']
            if xplan is not None and cov.dtype == torch.float64:
                # targets: one pass over each fold's condition averages; cross patients: list sums
                ctx.call('cpsd_colsum', pk.daddr(d_mu), B, Cm)
                self.gram_scatter(gram_c, pk.daddr(d_cov), B, Cm, Cm, cov, 3)
                lp, ll = ctypes_int_ptr(pk.iaddr(xplan['o_lptr'])), ctypes_int_ptr(pk.iaddr(xplan['o_list']))
                xs_ = self.ws('m_xsum', (nM, 128), torch.float64)
                ctx.call('cpsd_sum_mats_f64', ptr(None), ptr(ccs['G']), 128 * 128, lp, ll, 1.0,
                         ptr(cov, B * 128 * 128), 128 * 128, 128 * 128, nM)
                ctx.call('cpsd_sum_mats_f64', ptr(None), ptr(ccs['S']), 128, lp, ll, 1.0, ptr(xs_), 128, 128, nM)
                ctx.call('cpsd_cov_from_sums', ptr(cov, B * 128 * 128), 128, 128 * 128, ptr(xs_), 128,
                         ctypes_int_ptr(pk.iaddr(xplan['o_nrows'])), Cm, ptr(mu, s0 * Cm), Cm, ptr(None),
                         0, nM)
            else:
                ctx.call('cpsd_colsum', pk.daddr(d_mu), nS, Cm)
                self.gram_scatter(gram_c, pk.daddr(d_cov), nS, Cm, Cm, cov, 3)
            ncol = min(n_padC, R)
            if wplan is not None and cov.dtype == torch.float64:
                ip = lambda o: ctypes_int_ptr(pk.iaddr(o))
                n_dev = ip(o_cds)
                # a batch whose problems fit one wave of SMs and needs new bases anyway solves
                # everything cold in ONE launch (cold + warm would be two dependent launches of
                # which the second only halves its own latency); the bases serve later batches
                one_wave = wplan['ncold'] > 0 and nS <= self._sm_count()
                if one_wave:
                    self.eig_warm(cov, n_dev, 0, nS, vs['ev'], vs['evec'], out_idx=ip(wplan['o_out']))
                elif wplan['ncold']:
                    self.eig_warm(cov, n_dev, 0, wplan['ncold'], vs['ev'], vs['evec'],
                                  sel=ip(wplan['o_selc']), out_idx=ip(wplan['o_out']))
                if wplan['ncold']:
                    self.ortho_bases(vs['evec'], ip(wplan['o_cslot']), wplan['ncold'], vb['Q'], vb['Qf'],
                                     vb['n'])
                    vb['n'] += wplan['ncold']
                    vb['tab'] = wplan['tab']
                if wplan['nwarm'] and not one_wave:
                    self.rotate_sym(cov, vb['Q'], wplan['nwarm'], 'mv', sel=ip(wplan['o_selw']),
                                    base=ip(wplan['o_basew']))
                    self.eig_warm(cov, n_dev, 0, wplan['nwarm'], vs['ev'], vs['evec'],
                                  sel=ip(wplan['o_selw']), out_idx=ip(wplan['o_out']), V0=vb['Qf'],
                                  v0_idx=ip(wplan['o_v0']))
            else:
                self.eig_any(cov[:B], n_padC, ctypes_int_ptr(pk.iaddr(o_cds)), 0, B, 'mv', ncols=ncol,
                             out=(vs['ev'][:B], vs['evec'][:B]))
                if nM:
                    self.eig_any(cov[B:], n_padC, ctypes_int_ptr(pk.iaddr(o_cds + B)), 0, nM, 'mvx',
                                 ncols=ncol, out=(vs['ev'][s0:s0 + nM], vs['evec'][s0:s0 + nM]))
            vs['keys'].update(pending)
            vs['next'] += nM
            self.stats['view_solves'] = self.stats.get('view_solves', 0) + nS
            self.stats['view_problems'] = self.stats.get('view_problems', 0) + B * P
            if rank_done is not None:
                ctx.torch_stream.wait_event(rank_done)
            ctx.call('cpsd_mcca_mask_idx', ptr(vs['evec']), n_padC, n_padC * n_padC, ptr(vs['ev']),
                     n_padC, ptr(rank_dev) if use_rank else ptr(None), cdim_dev,
                     ctypes_int_ptr(pk.iaddr(o_slot)), R, Cm, ptr(Vr), ptr(d2), ptr(r_eff), B * P)
            ctx.call('cpsd_proj_nn', pk.daddr(d_pz), npz, max(Ks), T, R)
            ctx.call('cpsd_gram_tn', pk.daddr(d_g), ng, max(R, XR), max(R, XR))
            xc['keys'].update(xs_of)
            M = self.ws('m_M', (B, n_padM, n_padM))
            M.zero_()
            n_m = self.ws('m_nm', (B,), I32)
            cidx = self.ws('m_cidx', (B, P * R), I32)
            dh = self.ws('m_dh', (B, P * R))
            status = self.ws('m_status', (B,), I32)
            status.zero_()
            reg = -1.0 if self.regs is None else float(self.regs)
            if XR:
                ctx.call('cpsd_mcca_build_split', ptr(Gtx), P * R, R * P * R, ptr(Gxx), XR, XR * XR,
                         ctypes_int_ptr(pk.iaddr(o_xslot)), ptr(r_eff), P, R, reg, ptr(M), n_padM,
                         n_padM * n_padM, ptr(n_m), ptr(cidx), ptr(dh), Q, ptr(status), B)
            else:
                ctx.call('cpsd_mcca_build', ptr(Gtx), R, R * R, ptr(r_eff), P, R, reg, ptr(M), n_padM,
                         n_padM * n_padM, ptr(n_m), ptr(cidx), ptr(dh), Q, ptr(status), B)
            self.mark('mcca_gevp')
            evm, U = self.eig_any(M, n_padM, ptr(n_m), 0, B, 'mm', ncols=Q)
            ctx.call('cpsd_mcca_loadings', ptr(Vr), ptr(U), n_padM, n_padM * n_padM, ptr(None), 0,
                     ptr(r_eff), ptr(dh), P, R, Cm, Q, ptr(L), Q, B)
            if align_only:
                torch.cuda.synchronize(self.ctx.device)
                st = status.cpu().numpy()
                if st.any():
                    raise ValueError('MCCA: n_components=%d exceeds the total signal rank' % Q)
                return dict(loadings=L.view(B, P, Cm, Q).cpu().numpy(),
                            mu=self._slot_means(mu, slot, Cm), evals_mcca=evm[:, :Q].cpu().numpy(),
                            r_eff=r_eff.view(B, P).cpu().numpy(), shared=[s_.copy() for s_ in shared])
        # project every trial of every view into the pooled (trial x time*Q) matrix
        self.mark('project_pool')
        if tc_proj:
            tcp = self._tc_proj_ws(B * P, Q)
            ctx.call('cpsd_proj_tc_prep', ptr(L), Q, Cm * Q, ptr(None) if joint else ptr(mu),
                     ptr(None) if joint else ctypes_int_ptr(pk.iaddr(o_slot)),
                     Cm, cdim_dev, Q, tcp['ltc'], ptr(tcp['lthi']), ptr(tcp['ltlo']), ptr(tcp['mul']),
                     B * P)
            fbeg = np.searchsorted(rep, np.arange(J)).astype(np.int32)      # fold range of every replica
            fcnt = (np.searchsorted(rep, np.arange(J), side='right') - fbeg).astype(np.int32)
            ctx.call('cpsd_proj_tc_rep', ptr(tcp['xmaps']), ptr(tcp['ltmaps']), P, J, B, T, Q, tcp['ltc'],
                     ctypes.c_void_p(tcp['ntr'].ctypes.data), ctypes.c_void_p(tcp['nch'].ctypes.data),
                     ctypes.c_void_p(fbeg.ctypes.data), ctypes.c_void_p(fcnt.ctypes.data),
                     Nmax, ctypes_int_ptr(pk.iaddr(o_dst)), ptr(tcp['mul']), ptr(Zall), n_pad * F,
                     tcp['sms'])
        else:
            ctx.call('cpsd_proj_nn', ctypes_off(pk.daddr(d_pp), 0), B * P + B,
                     max(max(self.views[v].N for v in range(P)), n_te_max), T, Q)
        evals, k2_, St_, Ste, V, sweeps, kcap = yield from self._pooled_stage_run_gen(
            pk, d_p1, d_p2, B, Zall, pmu, Kall, n_pad, F, n_pool, n_te, o_npool, o_nall, o_ypool,
            n_te_max, want_details)
        self.mark('svm')
        yhat = self._decode(pk, d_svm, W, info, B, St_, Ste, k2, kcap, n_pad, o_ypool, n_pad, o_npool,
                            o_nte, n_te_max)
        self.mark('end')
        yield 'sync'
        self._verify_spec()
        yh = yhat.cpu().numpy()
        k2h = k2.cpu().numpy()
        st = status.cpu().numpy()
        self._check_decoder(info, B)
        if jvar:
            kjh = kj.cpu().numpy()
            if (kjh > Q).any():
                raise RuntimeError('JointPCA: a fold keeps %d components, above the batch layout of %d '
                                   'columns; pass joint_cap=%d' % (int(kjh.max()), Q, int(kjh.max()) + 4))
        if st.any():
            raise ValueError('MCCA: n_components=%d exceeds the total signal rank in fold(s) %s'
                             % (Q, np.nonzero(st)[0].tolist()))
        res = {'y_pred': [yh[f, :n_te[f]].copy() for f in range(B)], 'k2': k2h.tolist(),
               'h2d_bytes': pk.h2d_bytes, 'd2h_bytes': yh.nbytes + k2h.nbytes + st.nbytes}
        self.stats['launches_last_batch'] = ctx.launches() - launches0
        if want_details and joint:
            res['details'] = dict(
                loadings=L.view(B, P, Cm, Q).cpu().numpy(), evals_joint=evj[:, :Q].cpu().numpy(),
                pool_evals=evals.cpu().numpy(), svm_info=info.cpu().numpy(), W=_w_host(W),
                n_pool=list(n_pool), shared=[s.copy() for s in shared], lsq_status=stj.cpu().numpy(),
                k_joint=kj.cpu().numpy() if jvar else np.full(B, Q),
                bj_sweeps=None if sweeps is None else sweeps.cpu().numpy()[B:2 * B])
        elif want_details:
            res['details'] = dict(
                loadings=L.view(B, P, Cm, Q).cpu().numpy(), mu=self._slot_means(mu, slot, Cm),
                evals_mcca=evm[:, :Q].cpu().numpy(), r_eff=r_eff.view(B, P).cpu().numpy(),
                pool_evals=evals.cpu().numpy(), svm_info=info.cpu().numpy(),
                W=_w_host(W), n_pool=list(n_pool), shared=[s.copy() for s in shared],
                bj_sweeps=None if sweeps is None else sweeps.cpu().numpy()[B:2 * B])
        return res

    # ------------------------------------------------------------------ CCA / none batch
    def _batch_cca(self, batch, want_details):
        from .engine_cca import batch_cca
        return batch_cca(self, batch, want_details)


def plan_view_solves(code, tab, n_bases):
    """Warm-start plan of a batch of per-view eigenproblems.  ``code[i]`` = (replica, patient) pair
    of problem i, ``tab[pair]`` = index of the pair's basis or -1, ``n_bases`` = bases stored so far.
    The first problem of every pair without a basis is solved cold and becomes the pair's basis
    (bases are numbered in ascending pair order from ``n_bases``); all other problems are warm.
    Returns (cold problem indices, warm problem indices ascending, base_of[i] = basis of problem i
    or -1 for a cold one, the updated table -- a copy, committed by the caller once the cold
    solves are queued)."""
    code = np.asarray(code, dtype=np.int64)
    base_of = tab[code].astype(np.int32)
    miss = np.nonzero(base_of < 0)[0]
    cold = np.zeros(0, dtype=np.int64)
    if len(miss):
        uc, first_miss = np.unique(code[miss], return_index=True)
        cold = miss[first_miss]
        tab = tab.copy()
        tab[uc] = n_bases + np.arange(len(uc), dtype=np.int32)
        base_of = tab[code].astype(np.int32)
        base_of[cold] = -1
    warm_i = np.nonzero(base_of >= 0)[0]
    return cold, warm_i, base_of, tab


def _w_host(W):
    """Decoder state for the details dict: OvR weights, or the C-SVC's coef / rho / gamma."""
    if isinstance(W, dict):
        if W.get('bag'):
            return {}
        return {k: W[k].cpu().numpy() for k in ('coef', 'rho', 'gamma')}
    return W.cpu().numpy()


def _drain(g):
    """Runs a generator to completion and returns its value."""
    try:
        while True:
            next(g)
    except StopIteration as e:
        return e.value


def ctypes_int_ptr(address):
    return ctypes.c_void_p(address)


def ctypes_off(p, nbytes):
    return ctypes.c_void_p((p.value or 0) + nbytes)


def _p(x):
    """Tensor / None / c_void_p -> c_void_p."""
    return x if isinstance(x, ctypes.c_void_p) else ptr(x)
