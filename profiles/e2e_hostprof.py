"""cProfile of the host side of the streaming e2e path (cv_align_decode_stream, bench.py's e2e leg):
where the Python time of a 20-fold job goes."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cross_patient_speech_decoding_b200 import cv_align_decode_stream  # noqa: E402

pts = bench.make_data()
y0 = pts[0][1]
host_pts = [(torch.from_numpy(np.ascontiguousarray(X)).pin_memory(), y, ya) for X, y, ya in pts]
kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8, use_tensor_cores=True, max_batch=148)
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 21
group = int(sys.argv[2]) if len(sys.argv) > 2 else 7


def jobs(n, s0):
    for s in range(n):
        yield host_pts[0], host_pts[1:], bench.step_folds(y0, s0 + s)


for _ in cv_align_decode_stream(jobs(2 * depth, 77), depth=depth, group=group, **kw):
    pass
torch.cuda.synchronize()
n = 84
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in cv_align_decode_stream(jobs(n, 500), depth=depth, group=group, **kw):
    pass
torch.cuda.synchronize()
pr.disable()
dt = time.perf_counter() - t0
print('depth %d: %.2f ms per job (under cProfile), %.0f folds/s' % (depth, 1e3 * dt / n, 20 * n / dt))
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(28)
st.sort_stats('cumtime').print_stats(22)
