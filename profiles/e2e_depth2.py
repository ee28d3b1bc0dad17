"""e2e folds/s of cv_align_decode_stream as a function of the number of jobs in flight."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cross_patient_speech_decoding_b200 as cp  # noqa: E402

pts = bench.make_data()
y0 = pts[0][1]
host_pts = [(torch.from_numpy(np.ascontiguousarray(X)).pin_memory(), y, ya) for X, y, ya in pts]
kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8, use_tensor_cores=True, max_batch=20)


def jobs(n, s0):
    for s in range(n):
        yield host_pts[0], host_pts[1:], bench.step_folds(y0, s0 + s)


for spec in sys.argv[1:] or ['8:1', '14:7', '21:7']:
    depth, group = (int(x) for x in spec.split(':'))
    for _ in cp.cv_align_decode_stream(jobs(depth + 2, 77), depth=depth, group=group, **kw):
        pass
    torch.cuda.synchronize()
    n = 56
    cp.cv_align_decode_stream.idle_s = 0.0
    t0 = time.perf_counter()
    for _ in cp.cv_align_decode_stream(jobs(n, 500), depth=depth, group=group, **kw):
        pass
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('depth %2d group %d: %.2f ms per 20-fold job, %.0f folds/s, scheduler idle %.0f %%'
          % (depth, group, 1e3 * dt / n, 20 * n / dt, 100 * cp.cv_align_decode_stream.idle_s / dt), flush=True)
