"""e2e folds/s of cv_align_decode_stream vs number of steps in flight."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cross_patient_speech_decoding_b200 import cv_align_decode_stream  # noqa: E402

pts = bench.make_data()
host = [(torch.from_numpy(np.ascontiguousarray(X)).pin_memory(), y, ya) for X, y, ya in pts]
kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8, use_tensor_cores=True,
          max_batch=20)


def jobs(n, seed0):
    for s in range(n):
        yield host[0], host[1:], bench.step_folds(pts[0][1], seed0 + s)


for depth in (8, 8):
    for _ in cv_align_decode_stream(jobs(depth + 2, 10), depth=depth, **kw):
        pass
    torch.cuda.synchronize()
    cv_align_decode_stream.idle_s = 0.0
    t0 = time.perf_counter()
    n = 24
    for _ in cv_align_decode_stream(jobs(n, 100), depth=depth, **kw):
        pass
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('   host idle (all jobs waiting on GPU): %.1f ms/step' % (1e3 * cv_align_decode_stream.idle_s / n))
    st = torch.cuda.memory_stats()
    print('depth %d: %.1f ms/step, %.0f folds/s; reserved %.1f GB, device allocs %d, frees %d' % (
        depth, 1e3 * dt / n, 20 * n / dt, torch.cuda.memory_reserved() / 1e9, st['num_device_alloc'], st['num_device_free']))
