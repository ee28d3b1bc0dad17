"""Spatial averaging of electrodes over square contact regions (reference:
processing_utils/spatial_avg_subsampling.py:11-119).  Index generation on the host; the
averaging itself (``spatial_avg_data``) is a GPU gather-mean over the grid layout."""
import numpy as np

from .grid_subsampling import grid_susbsample_idxs


def spatial_avg_idxs(gridSize, contactSize):
    """Non-overlapping contactSize x contactSize windows, centred in the grid (:99-119)."""
    start = ((gridSize[0] % contactSize) // 2, (gridSize[1] % contactSize) // 2)
    return grid_susbsample_idxs(gridSize, (contactSize, contactSize),
                                (contactSize, contactSize), start)


def sig_regions(chanMap, contactSize, sigChan=None):
    """Array form of ``spatial_avg_sig_channels``: the averaging regions, optionally only those
    with at least one significant channel and fewer than half NaN electrodes (:30-69)."""
    chanMap = np.asarray(chanMap)
    if chanMap.shape[0] == 24:
        chanMap = chanMap[1:-1, :]
    elif chanMap.shape[1] == 24:
        chanMap = chanMap[:, 1:-1]
    regions = spatial_avg_idxs(chanMap.shape, contactSize)
    if sigChan is None:
        return regions
    sigChan = np.squeeze(np.asarray(sigChan))
    out = []
    for idxs in regions:
        elec = chanMap[idxs[:, 0], idxs[:, 1]]
        if np.sum(np.isnan(elec)) >= len(elec) / 2:
            continue
        good = ~np.isnan(elec)
        if np.intersect1d(sigChan, elec[good].astype(int)).size > 0:
            out.append(idxs[good])
    return out


def spatial_avg_sig_channels(pt, contactSize, dataPath, useSig=False):
    import scipy.io as sio
    chanMap = sio.loadmat(f'{dataPath}/{pt}/{pt}_channelMap.mat')['chanMap']
    sig = sio.loadmat(f'{dataPath}/{pt}/{pt}_sigChannel.mat')['sigChannel'] if useSig else None
    return sig_regions(chanMap, contactSize, sig)


def spatial_avg_data(data, avgIdxs):
    """(trials, grid_x, grid_y, time) -> (trials, time, regions): mean over each region's
    electrodes (:74-96), computed on the GPU."""
    from .device_subsample import spatial_average
    return spatial_average(data, avgIdxs)
