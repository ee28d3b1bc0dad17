"""CUDA-event timings of single kernels on headline-workload shapes (one B200).
Usage: python profiles/kernel_bench.py [eig] [svm] ...   (CPSD_LIB=<path> picks a library build)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cross_patient_speech_decoding_b200.device import Context, ptr  # noqa: E402

ctx = Context.get('cuda:0')
I32 = torch.int32


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def bench_eig(nprob=107, n=128):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((nprob, 400, n)) * np.linspace(3, 0.05, n)
    A = np.einsum('bkm,bkn->bmn', X, X)
    for f64 in (True, False):
        Ad = ctx.upload(A.astype(np.float64 if f64 else np.float32))
        work = torch.empty_like(Ad)
        ev = ctx.empty((nprob, n))
        V = ctx.empty((nprob, n, n))
        sw = ctx.zeros((nprob,), I32)
        name = 'cpsd_eig_sym_small_f64' if f64 else 'cpsd_eig_sym_small'
        for vecs in (True, False):
            def run():
                ctx.call(name, ptr(Ad), n, n * n, ptr(None), n, nprob, ptr(ev), n,
                         ptr(V) if vecs else ptr(None), n, n * n, 18, 1e-10 if f64 else 3e-7, ptr(sw))
            ms = timeit(run)
            print('eig_tile %s vecs=%d nprob=%d n=%d: %.3f ms  sweeps %s' % (
                'f64' if f64 else 'f32', vecs, nprob, n, ms, sw.cpu().numpy()[:4]))


if __name__ == '__main__':
    what = sys.argv[1:] or ['eig']
    if 'eig' in what:
        bench_eig()
