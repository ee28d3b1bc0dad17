import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cross_patient_speech_decoding_b200 import cv_align_decode_stream, engine
pts = bench.make_data()
host = [(torch.from_numpy(np.ascontiguousarray(X)).pin_memory(), y, ya) for X, y, ya in pts]
kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8, use_tensor_cores=True, max_batch=20)
def jobs(n, seed0):
    for s in range(n):
        yield host[0], host[1:], bench.step_folds(pts[0][1], seed0 + s)
for _ in cv_align_decode_stream(jobs(6, 10), depth=3, **kw): pass
torch.cuda.synchronize()
log = []
orig_view = engine.View.__init__
def view_init(self, *a, **k):
    t0 = time.perf_counter(); orig_view(self, *a, **k); log.append(('view', time.perf_counter() - t0))
engine.View.__init__ = view_init
orig_prep = engine.CVEngine._prepare_cross
def prep(self):
    t0 = time.perf_counter(); orig_prep(self); log.append(('prepare_cross', time.perf_counter() - t0))
engine.CVEngine._prepare_cross = prep
orig_init = engine.CVEngine._init
def init(self, *a, **k):
    t0 = time.perf_counter(); orig_init(self, *a, **k); log.append(('engine_init', time.perf_counter() - t0))
engine.CVEngine._init = init
t0 = time.perf_counter()
for _ in cv_align_decode_stream(jobs(12, 100), depth=3, **kw): pass
torch.cuda.synchronize()
print('total per job %.1f ms' % (1e3 * (time.perf_counter() - t0) / 12))
import collections
agg = collections.defaultdict(list)
for k, v in log: agg[k].append(1e3 * v)
for k, v in agg.items(): print(k, 'n=%d mean %.2f ms max %.2f' % (len(v), np.mean(v), np.max(v)), [round(x, 1) for x in v[:10]])
