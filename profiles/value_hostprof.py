"""cProfile of the host side of the resident-data path (bench.py's `value` leg): CVEngine.run over
48 steps = 960 folds in batches of 137 on two lanes."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cross_patient_speech_decoding_b200.engine import CVEngine  # noqa: E402

pts = bench.make_data()
y0 = pts[0][1]
eng = CVEngine(pts[0], pts[1:], method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8,
               use_tensor_cores=True, max_batch=148)


def folds(s0, n):
    out = []
    for s in range(n):
        out += bench.step_folds(y0, s0 + s)
    return out


eng.run(folds(0, 15))
eng.run(folds(100, 15))
torch.cuda.synchronize()
fl = folds(1000, 48)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
eng.run(fl)
torch.cuda.synchronize()
pr.disable()
dt = time.perf_counter() - t0
print('%.1f ms for %d folds -> %.0f folds/s (under cProfile); host_pack_ms %.1f' % (1e3 * dt, len(fl), len(fl) / dt, eng.stats.get('host_pack_ms', 0)))
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(30)

# which named workspaces are (re)allocated inside a steady-state run
from cross_patient_speech_decoding_b200 import engine as E  # noqa: E402
log = []
orig = E.CVEngine.ws


def ws(self, name, shape, dtype=E.F32):
    import numpy as np
    n = int(np.prod(shape))
    t = self._ws.get(name)
    if t is None or t.numel() < n or t.dtype != dtype:
        log.append((self.lane, name, tuple(shape), str(dtype), None if t is None else t.numel()))
    return orig(self, name, shape, dtype)


E.CVEngine.ws = ws
eng.run(folds(2000, 48))
torch.cuda.synchronize()
print('workspace allocations in a steady-state run:', len(log))
for r in log[:40]:
    print('  ', r)
