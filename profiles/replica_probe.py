"""GPU time of one grouped engine (7 replicas x 20 folds): constructor (uploads, casts, fold-invariant
statistics of 7 x 7 cross patients) and the 140-fold batch with stage timers."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cross_patient_speech_decoding_b200.engine import CVEngine  # noqa: E402

pts = bench.make_data()
y0 = pts[0][1]
host = [(torch.from_numpy(np.ascontiguousarray(X)).pin_memory(), y, ya) for X, y, ya in pts]
J = int(sys.argv[1]) if len(sys.argv) > 1 else 7
kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8, use_tensor_cores=True, max_batch=148)
for trial in range(3):
    folds, rep = [], []
    for j in range(J):
        f = bench.step_folds(y0, 100 * trial + j)
        folds += f
        rep += [j] * len(f)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng = CVEngine(host[0], host[1:], replicas=[(host[0], host[1:])] * (J - 1), **kw)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    eng.profile = True
    res = eng.run(folds, rep=rep if J > 1 else None)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print('J=%d: constructor %.1f ms (uploads %.0f MB), run %.1f ms -> %.0f folds/s incl. constructor; stages %s'
          % (J, 1e3 * (t1 - t0), J * sum(v.h2d_bytes for v in eng.views) / 1e6, 1e3 * (t2 - t1),
             len(folds) / (t2 - t0), {k: round(v, 2) for k, v in eng.collect_marks().items()}), flush=True)
    del eng
