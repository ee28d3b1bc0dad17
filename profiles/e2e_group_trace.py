"""Host / device timeline of the grouped streaming path (7 replicas per engine, 2-3 engines in
flight): when each group is submitted, when its uploads and its batch finish on the device."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cross_patient_speech_decoding_b200 as cp  # noqa: E402
from cross_patient_speech_decoding_b200 import engine as E  # noqa: E402

pts = bench.make_data()
y0 = pts[0][1]
host = [(torch.from_numpy(np.ascontiguousarray(X)).pin_memory(), y, ya) for X, y, ya in pts]
kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8, use_tensor_cores=True, max_batch=20)
depth, group = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (14, 7)
base = torch.cuda.Event(enable_timing=True)
log = []
orig_init = E.CVEngine.__init__
orig_run = E.CVEngine.run_gen


def init(self, *a, **k):
    t0 = time.perf_counter()
    orig_init(self, *a, **k)
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(self.stream)
    self._trace = dict(t_sub0=t0, t_sub1=time.perf_counter(), ev_ctor=ev, lane=self.lane)
    log.append(self._trace)


def run_gen(self, *a, **k):
    out = yield from orig_run(self, *a, **k)
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(self.stream)
    self._trace.update(t_done=time.perf_counter(), ev_done=ev)
    return out


E.CVEngine.__init__ = init
E.CVEngine.run_gen = run_gen


def jobs(n, s0):
    for s in range(n):
        yield host[0], host[1:], bench.step_folds(y0, s0 + s)


for _ in cp.cv_align_decode_stream(jobs(2 * depth, 77), depth=depth, group=group, **kw):
    pass
torch.cuda.synchronize()
del log[:]
base.record()
T0 = time.perf_counter()
n = 6 * group
for _ in cp.cv_align_decode_stream(jobs(n, 500), depth=depth, group=group, **kw):
    pass
torch.cuda.synchronize()
dt = time.perf_counter() - T0
print('depth %d group %d: %.0f folds/s' % (depth, group, 20 * n / dt))
for g in log:
    print('lane %2d  submit %6.1f..%6.1f ms (host)   ctor work done on device %6.1f ms   batch done on device %6.1f ms   result on host %6.1f ms'
          % (g['lane'], 1e3 * (g['t_sub0'] - T0), 1e3 * (g['t_sub1'] - T0), base.elapsed_time(g['ev_ctor']),
             base.elapsed_time(g['ev_done']), 1e3 * (g['t_done'] - T0)))
