"""Wall-clock breakdown of one cv_align_decode call (host float64 buffers -> predictions)."""
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
host = [(torch.from_numpy(np.ascontiguousarray(X)).pin_memory(), y, ya) for X, y, ya in pts]
kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8, use_tensor_cores=True)
for it in range(4):
    folds = bench.step_folds(pts[0][1], 900 + it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng = CVEngine(host[0], host[1:], **kw)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    eng.profile = True
    res = eng.run(folds)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    st = eng.collect_marks()
    print('call %d: engine build %.1f ms, run %.1f ms (host pack %.1f ms), stages %s' % (
        it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), eng.stats.get('host_pack_ms', 0),
        {k: round(v, 2) for k, v in st.items()}))
