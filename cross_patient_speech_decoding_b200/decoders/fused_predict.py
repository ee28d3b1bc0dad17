"""Low-latency predict for a fitted cross-patient decoder (BASELINE config 5, the realtime
path): ``FusedPredictor(model)`` captures the three fitted stages that
``crossPtDecoder.predict`` chains (decoders/cross_pt_decoders.py:70-71) -- the target's aligner
transform ``(X - mu) A``, the ``DimRedReshape`` PCA transform and the one-vs-rest linear
decision -- as device tensors, and ``predict`` runs them as one kernel
(``cpsd_predict_fused``): one upload of the trials, one launch, one read-back of the labels.
The decoder must be ``make_pipeline(DimRedReshape(PCA, ...), LinearSVC())`` from this package,
or the same pipeline ending in this package's ``SVC`` (then the kernel stops at the PCA scores
and ``cpsd_svc_predict_ovo`` casts the one-vs-one votes: two launches)."""
import numpy as np
import torch

from ..device import Context, ptr


def _target_map(model):
    """(mu (C,) or None, A (C, Q)) of the target patient's transform for each decoder class."""
    name = type(model).__name__
    if name == 'crossPtDecoder_mcca':
        mc = model.aligner.mcca
        return np.asarray(mc.means_[0]), np.asarray(mc.loadings_[0])
    if name == 'crossPtDecoder_jointDimRed':
        return None, np.asarray(model.joint_dr.transforms[0])
    if name in ('crossPtDecoder_sepAlign', 'crossPtDecoder_sepDimRed'):
        comps = np.asarray(model.tar_dr.components_)             # (d, C)
        if name == 'crossPtDecoder_sepDimRed':
            comps = comps[:model.common_dim]
        return np.asarray(model.tar_dr.mean_), comps.T
    raise TypeError('unsupported decoder class %s' % name)


class FusedPredictor:
    def __init__(self, model, device=None):
        steps = model.decoder.steps
        pca, svm = steps[0][1].transformer, steps[-1][1]
        self.svc = getattr(svm, '_model', None)               # svm.SVC: libsvm-style C-SVC, one-vs-one
        if (not hasattr(svm, '_W') and self.svc is None) or len(steps) != 2:
            raise TypeError('FusedPredictor needs make_pipeline(DimRedReshape(PCA), LinearSVC() or SVC())')
        mu, A = _target_map(model)
        self.ctx = ctx = Context.get(device)
        self.C, self.Q = (int(v) for v in A.shape)
        comps = np.asarray(pca.components_)                       # (k2, F)
        self.k2, self.F = (int(v) for v in comps.shape)
        if self.F % self.Q:
            raise ValueError('PCA feature count is not a multiple of the latent size')
        self.T = self.F // self.Q
        self.classes_ = np.asarray(svm.classes_)
        self.mu = None if mu is None else ctx.upload(mu, np.float32)
        self.A = ctx.upload(A, np.float32)
        self.pmean = ctx.upload(np.asarray(pca.mean_), np.float32)
        self.P = ctx.upload(np.ascontiguousarray(comps.T), np.float32)
        if self.svc is None:
            self.W = ctx.upload(np.asarray(svm._W), np.float64)
            self.ncls = int(svm._W.shape[0])
        else:
            if self.svc['k'] != self.k2:
                raise ValueError('SVC was fitted on a different number of PCA components')
            self.W, self.ncls = None, len(self.classes_)
        self.cls = ctx.upload(self.classes_.astype(np.int32), np.int32)
        self._cap = 0

    def _buffers(self, n):
        if n > self._cap:
            self._cap = max(n, 2 * self._cap, 16)
            self._hx = torch.empty((self._cap, self.T, self.C), dtype=torch.float64).pin_memory()
            self._dx = self.ctx.empty((self._cap, self.T, self.C), torch.float64)
            self._yh = self.ctx.empty((self._cap,), torch.int32)
            self._dec = self.ctx.empty((self._cap, self.ncls), torch.float64)
            self._hy = torch.empty((self._cap,), dtype=torch.int32).pin_memory()
            k2p = (self.k2 + 31) // 32 * 32
            self._wsp = self.ctx.empty((self._cap * min(self.T, 16) * k2p,), torch.float64)
            self._wsc = self.ctx.zeros((self._cap,), torch.int32)
            self._sc = self.ctx.empty((self.k2, self._cap)) if self.svc is not None else None

    def predict(self, X, return_decision=False):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim != 3 or X.shape[1] != self.T or X.shape[2] != self.C:
            raise ValueError('expected trials of shape (n, %d, %d)' % (self.T, self.C))
        n = X.shape[0]
        self._buffers(n)
        self._hx[:n].copy_(torch.from_numpy(np.ascontiguousarray(X)))
        # few trials: split every trial over several CTAs (time slices) to fill the SMs
        nsplit = max(1, min(self.T, 16, 592 // n))
        self._dx[:n].copy_(self._hx[:n], non_blocking=True)
        self.ctx.call('cpsd_predict_fused', ptr(self._dx), n, self.T, self.C, ptr(self.mu), ptr(self.A),
                      self.Q, ptr(self.pmean), ptr(self.P), self.k2, ptr(self.W), ptr(self.cls),
                      self.ncls, ptr(self._yh), ptr(self._dec), nsplit, ptr(self._wsp), ptr(self._wsc),
                      ptr(self._sc), self._cap)
        if self.svc is not None:
            if return_decision:
                raise ValueError('return_decision is only available with the linear decoder')
            m = self.svc
            self.ctx.call('cpsd_svc_predict_ovo', ptr(m['St']), m['lds'], 0, ptr(self._sc), self._cap, 0,
                          ptr(None), self.k2, ptr(None), m['n'], m['n'], ptr(None), n, ptr(m['y']), 0,
                          ptr(m['classes_dev']), self.ncls, m['kernel'], ptr(m['gamma_dev']),
                          ptr(m['coef_dev']), m['lds'], ptr(m['rho_dev']), ptr(self._yh), ptr(None),
                          self.k2, 1)
        self._hy[:n].copy_(self._yh[:n], non_blocking=True)
        torch.cuda.current_stream(self.ctx.device).synchronize()
        out = self._hy[:n].numpy().astype(self.classes_.dtype)
        if return_decision:
            return out, self._dec[:n].cpu().numpy()
        return out
