"""cProfile of the host-side packing of one MCCA batch (index tables + descriptors)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cross_patient_speech_decoding_b200.engine import CVEngine, _drain  # noqa: E402

pts = bench.make_data()
eng = CVEngine(pts[0], pts[1:], method='mcca', n_comp=30, regs=0.5, pca_var=0.8,
               use_tensor_cores=True, max_batch=138)
folds = []
while len(folds) < 138:
    folds += bench.step_folds(pts[0][1], 5000 + len(folds))
folds = folds[:138]
eng.run(folds)
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    eng._ensure_ready()
    _drain(eng._mcca_start(folds, False))
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(30)
