"""Is the streaming e2e path bound by the host, the PCIe upload or the GPU?  Runs the same 20-fold
jobs (a) from pinned host float64 buffers, (b) from device-resident float64 tensors, both 8 deep,
and prints the per-job wall time and the share of it the scheduler spent idle (all jobs waiting
on the GPU)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import cross_patient_speech_decoding_b200 as cp
pts = bench.make_data()
host = [(torch.from_numpy(np.ascontiguousarray(X)).pin_memory(), y, ya) for X, y, ya in pts]
dev = [(X.cuda(), y, ya) for X, y, ya in host]
kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8, use_tensor_cores=True, max_batch=20)
def jobs(src, n, seed0):
    for s in range(n):
        yield src[0], src[1:], bench.step_folds(pts[0][1], seed0 + s)
for depth in (8, 16):
    for name, src in (('host', host), ('device', dev)):
        for _ in cp.cv_align_decode_stream(jobs(src, depth + 2, 10), depth=depth, **kw): pass
        torch.cuda.synchronize()
        cp.cv_align_decode_stream.idle_s = 0.0
        n = 40
        t0 = time.perf_counter()
        for _ in cp.cv_align_decode_stream(jobs(src, n, 100), depth=depth, **kw): pass
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print('%s inputs, depth %d: %.2f ms per 20-fold job (%.0f folds/s), scheduler idle %.0f %%' % (
            name, depth, 1e3 * dt / n, 20 * n / dt, 100 * cp.cv_align_decode_stream.idle_s / dt), flush=True)
