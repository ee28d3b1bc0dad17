"""Sliding-window sub-grids of an electrode array (reference:
processing_utils/grid_subsampling.py:8-98).  ``grid_susbsample_idxs`` keeps the reference's
spelling; the window order is that of ``np.meshgrid(startX, startY)`` flattened, i.e. the
column start varies slowest, and inside a window the column index varies slowest too."""
import numpy as np


def grid_susbsample_idxs(gridSize, winSize, step=(1, 1), start=(0, 0)):
    """All (row, col) index arrays, one ``(winSize[0]*winSize[1], 2)`` array per window."""
    sx = np.arange(start[0], gridSize[0] - winSize[0] + 1, step[0])
    sy = np.arange(start[1], gridSize[1] - winSize[1] + 1, step[1])
    # offsets inside one window, same traversal as the window starts
    oy, ox = np.meshgrid(np.arange(winSize[1]), np.arange(winSize[0]), indexing='ij')
    off = np.stack([ox.ravel(), oy.ravel()], axis=1)
    out = []
    for y0 in sy:
        for x0 in sx:
            out.append(off + np.array([x0, y0]))
    return out


def trim_channel_map(chanMap, winSize=None):
    """Drops the all-NaN border of the 24-wide maps (grid_subsampling.py:33-38); the window is
    transposed when the long axis comes first."""
    chanMap = np.asarray(chanMap)
    if chanMap.shape[0] == 24:
        chanMap = chanMap[1:-1, :]
        if winSize is not None:
            winSize = (winSize[1], winSize[0])
    elif chanMap.shape[1] == 24:
        chanMap = chanMap[:, 1:-1]
    return chanMap, winSize


def sig_channels_in_windows(chanMap, sigChan, winSize, step=(1, 1)):
    """Array form of ``grid_subsample_sig_channels``: for every window the positions (inside
    the sorted significant-channel list) of the significant channels it covers; windows
    without any are dropped (grid_subsampling.py:40-61)."""
    chanMap, winSize = trim_channel_map(chanMap, winSize)
    sigChan = np.squeeze(np.asarray(sigChan))
    out = []
    for idxs in grid_susbsample_idxs(chanMap.shape, winSize, step=step):
        elec = chanMap[idxs[:, 0], idxs[:, 1]]
        elec = elec[~np.isnan(elec)].astype(int)
        _, sig_idx, _ = np.intersect1d(sigChan, elec, return_indices=True)
        if len(sig_idx) > 0:
            out.append(sig_idx)
    return out


def grid_subsample_sig_channels(pt, winSize, dataPath, step=(1, 1)):
    """File-based form with the reference's signature (loads ``<pt>_channelMap.mat`` and
    ``<pt>_sigChannel.mat`` like grid_subsampling.py:26-31)."""
    import scipy.io as sio
    chanMap = sio.loadmat(f'{dataPath}/{pt}/{pt}_channelMap.mat')['chanMap']
    sigChan = sio.loadmat(f'{dataPath}/{pt}/{pt}_sigChannel.mat')['sigChannel']
    return sig_channels_in_windows(chanMap, sigChan, winSize, step=step)
