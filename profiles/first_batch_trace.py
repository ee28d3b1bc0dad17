"""Kernel table (torch.profiler / CUPTI) of a fresh grouped engine: the constructor and the first --
in the end-to-end path the only -- batch, separately.  Usage: first_batch_trace.py [replicas]"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cross_patient_speech_decoding_b200.engine import CVEngine  # noqa: E402

pts = bench.make_data()
y0 = pts[0][1]
host = [(torch.from_numpy(np.ascontiguousarray(X)).pin_memory(), y, ya) for X, y, ya in pts]
J = int(sys.argv[1]) if len(sys.argv) > 1 else 7
kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8, use_tensor_cores=True, max_batch=148)


def table(prof, title):
    rows = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.replace('void ', '').replace('(anonymous namespace)::', '').split('(')[0][:60]
            r = rows.setdefault(name, [0, 0.0])
            r[0] += 1
            r[1] += ev.device_time_total if hasattr(ev, 'device_time_total') else ev.cuda_time_total
    tot = sum(r[1] for r in rows.values())
    print('%s: %.2f ms of kernels / copies' % (title, tot / 1e3))
    print('  k_eig_tile launches (ms, in order):',
          [(ev.name.split('<')[1][:6], round((ev.device_time_total if hasattr(ev, 'device_time_total')
                                              else ev.cuda_time_total) / 1e3, 2))
           for ev in sorted(prof.events(), key=lambda e: e.time_range.start)
           if ev.device_type == torch.autograd.DeviceType.CUDA and 'k_eig_tile' in ev.name])
    for name, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:24]:
        print('  %-60s x%-4d %8.3f ms' % (name, n, us / 1e3))


for trial in range(2):
    folds, rep = [], []
    for j in range(J):
        f = bench.step_folds(y0, 100 * trial + j)
        folds += f
        rep += [j] * len(f)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as p1:
        eng = CVEngine(host[0], host[1:], replicas=[(host[0], host[1:])] * (J - 1), **kw)
        torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as p2:
        res = eng.run(folds, rep=rep if J > 1 else None)
        torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as p3:
        res = eng.run(folds, rep=rep if J > 1 else None)
        torch.cuda.synchronize()
    if trial == 1:
        table(p3, 'second batch (%d folds)' % len(folds))
        table(p1, 'constructor (J=%d)' % J)
        table(p2, 'first batch (%d folds)' % len(folds))
    del eng
