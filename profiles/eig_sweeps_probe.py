"""Sweep counts of every fp64 tile-Jacobi launch of a fresh engine's first batch (by call tag)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cross_patient_speech_decoding_b200 import engine as E  # noqa: E402
from cross_patient_speech_decoding_b200.device import ptr  # noqa: E402

pts = bench.make_data()
y0 = pts[0][1]
host = [(torch.from_numpy(np.ascontiguousarray(X)).pin_memory(), y, ya) for X, y, ya in pts]
kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8, use_tensor_cores=True, max_batch=148)
log = []
orig = E.CVEngine.eig_any


def eig_any(self, A, n_pad, n_dev, n_fixed, nprob, tag, ncols=None, vecs=True, out=None):
    if A.dtype == torch.float64 and n_pad <= 128 and nprob:
        evals, evecs = out if out is not None else (self.ws(tag + '_ev', (nprob, n_pad)),
                                                    self.ws(tag + '_evec', (nprob, n_pad, n_pad)))
        sw = torch.zeros(nprob, dtype=torch.int32, device=A.device)
        dg = torch.diagonal(A, dim1=1, dim2=2)
        rng = (dg.abs().amax(1) / dg.abs().clamp_min(1e-300).amin(1)).max().item()
        self.ctx.call('cpsd_eig_sym_small_f64', ptr(A), n_pad, n_pad * n_pad, E._p(n_dev), n_fixed, nprob,
                      ptr(evals), n_pad, ptr(evecs if vecs else None), n_pad, n_pad * n_pad,
                      self.eig_sweeps + 6, 1e-10, ptr(sw))
        log.append((tag, nprob, vecs, sw, evals, rng))
        return evals, evecs
    return orig(self, A, n_pad, n_dev, n_fixed, nprob, tag, ncols=ncols, vecs=vecs, out=out)


E.CVEngine.eig_any = eig_any
eng = E.CVEngine(host[0], host[1:], **kw)
print('eig_sweeps', eng.eig_sweeps)
eng.run(bench.step_folds(y0, 0))
torch.cuda.synchronize()
for tag, nprob, vecs, sw, ev, rng in log:
    s = sw.cpu().numpy()
    e = ev.cpu().numpy()
    print('%-5s nprob %3d vecs %d sweeps min %d max %d  diag range %.1e  ev[0] %.3e ev[-1] %.3e' % (
        tag, nprob, vecs, s.min(), s.max(), rng, e[:, 0].max(), e[:, :128].min()))
