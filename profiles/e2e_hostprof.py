import cProfile, os, pstats, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cross_patient_speech_decoding_b200 import cv_align_decode_stream
pts = bench.make_data()
host = [(torch.from_numpy(np.ascontiguousarray(X)).pin_memory(), y, ya) for X, y, ya in pts]
kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8, use_tensor_cores=True, max_batch=20)
def jobs(n, seed0):
    for s in range(n):
        yield host[0], host[1:], bench.step_folds(pts[0][1], seed0 + s)
for _ in cv_align_decode_stream(jobs(6, 10), depth=4, **kw): pass
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in cv_align_decode_stream(jobs(24, 100), depth=4, **kw): pass
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(25)
