"""CUDA-event time of the pooled-Gram call (k_split_tf32_batched + k_gram_tc) on the headline shape:
138 problems of 1152 x 6000.  CPSD_LIB=<path> picks a library build (experiment variants)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cross_patient_speech_decoding_b200 import _lib  # noqa: E402
from cross_patient_speech_decoding_b200.device import Context, addr, ptr  # noqa: E402

ctx = Context.get('cuda:0')
nprob, n, k = 138, 1152, 6000
Z = torch.randn((nprob, n, k), device='cuda')
out = ctx.zeros((nprob, n, n))
rec = np.zeros(nprob, dtype=_lib.GRAM_NT_DESC)
for p in range(nprob):
    rec[p] = (addr(Z, p * n * k), addr(Z, p * n * k), addr(out, p * n * n), n, n, k, k, k, n, 1, 1.0)
nbytes = int(ctx.lib.cpsd_gram_nt_tc_ws_bytes(nprob))
split = ctx.empty((2 * nprob * n * k,))
maps = ctx.empty((nbytes + 64,), torch.uint8)
stage = torch.empty((nbytes + 64,), dtype=torch.uint8).pin_memory()
a = (maps.data_ptr() + 63) & ~63


def run():
    ctx.call('cpsd_gram_nt_tc', ctypes.c_void_p(rec.ctypes.data), nprob, n, n, ptr(split), split.numel(),
             ctypes.c_void_p(a), ctypes.c_void_p(stage.data_ptr()))


run()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
flops = 2.0 * n * n * k * nprob
ref = (Z[0, :128].double() @ Z[0, :256].double().T).float()
err = ((out[0, :128, :256] - ref).abs().max() / ref.abs().max()).item()
print('%s: split + gram %.3f ms (%.0f algorithmic TFLOP/s), rel err of a tile %.2e'
      % (os.environ.get('CPSD_LIB', 'default'), ms, flops / ms / 1e9, err))
