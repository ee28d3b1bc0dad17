"""GPU side of the electrode-subsampling analyses: the trials of every patient are uploaded
once (``resident``), each subsample is a channel gather on the device (``gather_channels``)
whose result goes straight into ``cv_align_decode`` / ``cv_align_decode_stream`` /
``CVEngine`` -- no host copy of the 50 x sliced arrays the reference materialises
(scripts/aligned_decode_grid_subsample.py:280-300)."""
import numpy as np
import torch

from ..device import Context, ptr

I32 = torch.int32


def _mark_ready(t):
    """Tags a freshly produced CUDA tensor with an event on the producing stream; the engine's
    lane stream waits on it before its first kernel reads the tensor (engine.View)."""
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(t.device))
    t._cpsd_ready = ev
    return t


def resident(X, device=None):
    """Host (trials, time, channels) array -> fp32 CUDA tensor (uploaded once)."""
    ctx = Context.get(device)
    X = np.ascontiguousarray(X)
    t = torch.from_numpy(X)
    t = t.pin_memory()
    raw = t.to(ctx.device, non_blocking=True)
    if raw.dtype == torch.float64:
        out = ctx.empty(tuple(X.shape))
        ctx.call('cpsd_cast_f64_f32', ptr(raw), ptr(out), raw.numel())
        out._cpsd_src = (t, raw)         # pinned source + raw copy stay alive with the tensor
        return _mark_ready(out)
    out = raw.to(torch.float32)
    out._cpsd_src = (t, raw)
    return _mark_ready(out)


def gather_channels(X_dev, idx, device=None):
    """X_dev (trials, time, channels) fp32 CUDA tensor -> X_dev[:, :, idx] (new CUDA tensor)."""
    ctx = Context.get(device if device is not None else X_dev.device)
    assert X_dev.is_cuda and X_dev.dtype == torch.float32 and X_dev.dim() == 3
    ev = getattr(X_dev, '_cpsd_ready', None)
    if ev is not None:          # the source may still be uploading on another stream
        torch.cuda.current_stream(X_dev.device).wait_event(ev)
    X_dev = X_dev.contiguous()
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    assert idx.ndim == 1 and idx.size > 0 and idx.min() >= 0 and idx.max() < X_dev.shape[2]
    idx_d = ctx.upload(idx, np.int32)
    N, T, C = (int(v) for v in X_dev.shape)
    out = ctx.empty((N, T, idx.size))
    ctx.call('cpsd_gather_channels', ptr(X_dev), C, ptr(idx_d), int(idx.size), ptr(out), int(idx.size),
             N * T)
    out._cpsd_src = (X_dev, idx_d)
    return _mark_ready(out)


def gather_trials(X_dev, idx, device=None):
    """X_dev (trials, time, channels) fp32 CUDA tensor -> X_dev[idx] (new CUDA tensor): the
    ``x[samp_idx]`` of scripts/aligned_decode_cross_patient_subsample.py:309-311 on the device."""
    ctx = Context.get(device if device is not None else X_dev.device)
    assert X_dev.is_cuda and X_dev.dtype == torch.float32 and X_dev.dim() == 3
    ev = getattr(X_dev, '_cpsd_ready', None)
    if ev is not None:
        torch.cuda.current_stream(X_dev.device).wait_event(ev)
    X_dev = X_dev.contiguous()
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    assert idx.ndim == 1 and idx.size > 0 and idx.min() >= 0 and idx.max() < X_dev.shape[0]
    idx_d = ctx.upload(idx, np.int32)
    N, T, C = (int(v) for v in X_dev.shape)
    out = ctx.empty((idx.size, T, C))
    ctx.call('cpsd_gather_trials', ptr(X_dev), T * C, ptr(idx_d), int(idx.size), ptr(out))
    out._cpsd_src = (X_dev, idx_d)
    return _mark_ready(out)


def spatial_average(data, avgIdxs, device=None):
    """spatial_avg_data: data (trials, grid_x, grid_y, time) host float array, avgIdxs list of
    (n_i, 2) grid index arrays -> (trials, time, regions) float64 numpy array."""
    ctx = Context.get(device)
    data = np.ascontiguousarray(data, dtype=np.float64)
    ntr, gx, gy, T = data.shape
    reg_ptr = np.concatenate([[0], np.cumsum([len(ix) for ix in avgIdxs])]).astype(np.int32)
    flat = np.concatenate([np.asarray(ix)[:, 0] * gy + np.asarray(ix)[:, 1] for ix in avgIdxs])
    d = ctx.upload(data.reshape(ntr, gx * gy, T))
    rp, re = ctx.upload(reg_ptr, np.int32), ctx.upload(flat, np.int32)
    out = ctx.empty((ntr, T, len(avgIdxs)), torch.float64)
    ctx.call('cpsd_region_mean_f64', ptr(d), ntr, gx * gy, T, ptr(rp), ptr(re), len(avgIdxs), ptr(out))
    return out.cpu().numpy()
