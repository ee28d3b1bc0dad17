"""Poisson-disk electrode subsampling at a given pitch (reference:
processing_utils/poisson_disk_sampling.py:9-176, Bridson 2007 dart throwing on a background
grid).  The numpy global RNG is consumed in exactly the reference's order --
``choice(avail, n, replace=False)`` then ``rand(n, ndim)`` per round, one final ``choice`` when
too many points were accepted -- so a seeded run selects the same electrodes."""
import numpy as np


def knn_search(pts, newPts, k):
    """Brute-force k nearest neighbours: (indices, distances), ascending
    (poisson_disk_sampling.py:154-176); all queries at once instead of a Python loop."""
    pts, newPts = np.asarray(pts, dtype=float), np.asarray(newPts, dtype=float)
    d = np.sqrt(np.sum((pts[None, :, :] - newPts[:, None, :]) ** 2, axis=2))
    order = np.argsort(d, axis=1)[:, :k]
    return order, np.sort(d, axis=1)[:, :k]


def min_neighbor_distance(pts, newPts):
    """Distance of every new point to its nearest OTHER point (the query points are part of
    ``pts``, so the nearest hit is the point itself; poisson_disk_sampling.py:137-151)."""
    return knn_search(pts, newPts, 2)[1][:, 1]


def poisson_disk_sampling(domain, spacing, nPoints, threshold=60, showIter=False, maxIter=1000):
    domain = tuple(domain)
    ndim = len(domain)
    cell = spacing / np.sqrt(ndim)
    axes = np.meshgrid(*[np.arange(1, s + 1, cell) for s in domain], indexing='ij')
    shape = axes[0].shape
    corners = np.column_stack([a.ravel() for a in axes])
    empty = np.ones(corners.shape[0], dtype=bool)
    n_empty = int(empty.sum())
    score = np.zeros(corners.shape[0], dtype=int)
    kept = []
    n_kept = 0
    it = 0
    while n_kept < nPoints and n_empty > 0:
        if it > maxIter:
            print(f'Reached max iterations with {n_kept} points. Trying sampling again.')
            return poisson_disk_sampling(domain, spacing, nPoints, threshold)
        avail = np.where(empty)[0]
        n_throw = np.minimum(n_empty, nPoints)
        cells = np.random.choice(avail, n_throw, replace=False)
        darts = corners[cells] + cell * np.random.rand(n_throw, ndim)
        everything = np.vstack((kept, darts)) if len(kept) > 0 else darts
        ok = np.all(darts < domain, axis=1) & (min_neighbor_distance(everything, darts) > spacing)
        missed, darts = darts[~ok, :], darts[ok, :]
        hit = np.floor((darts + cell - 1) / cell).astype(int)
        empty[np.ravel_multi_index(hit.T - 1, shape)] = False
        bad = np.floor((missed + cell - 1) / cell).astype(int)
        score[np.ravel_multi_index(bad.T - 1, shape)] += 1     # (repeated cells count once)
        empty &= score < threshold
        n_empty = int(empty.sum())
        kept.extend(darts)
        n_kept += darts.shape[0]
        it += 1
        if showIter:
            print(f'Iteration: {it}    Points Created: {n_kept}   EmptyGrid: {n_empty}')
    pts = np.vstack(kept)
    if n_kept > nPoints:
        pts = pts[np.random.choice(pts.shape[0], nPoints, replace=False)]
    return pts


# array dimensions / electrode counts of the two array types (poisson_disk_sampling.py:36-43)
_ARRAYS = {('S14', 'S22', 'S23', 'S26'): (11.3, 22.5, 128),
           ('S33', 'S39', 'S58', 'S62'): (37.8, 20.6, 256)}


def pitch_sig_channels(chanMap, sigChan, pitch, mmX, mmY, maxElec):
    """Array form of ``pitch_subsample_sig_channels`` (poisson_disk_sampling.py:28-77)."""
    chanMap = np.asarray(chanMap)
    if chanMap.shape[1] == 24:
        chanMap = chanMap[:, 1:-1]
    sigChan = np.squeeze(np.asarray(sigChan))
    nElec = round(mmX * mmY / pitch ** 2)
    if nElec >= maxElec:
        elec = np.arange(1, maxElec + 1)
    else:
        gx, gy = chanMap.shape
        spacing = np.floor(np.sqrt(gx * gy / nElec))
        idx = np.round(poisson_disk_sampling((gx, gy), spacing, nElec)).astype(int) - 1
        elec = np.nan_to_num(chanMap[idx[:, 0], idx[:, 1]], nan=-1).astype(int)
        if elec.shape[0] < nElec and spacing == 1:
            rest = np.setdiff1d(np.arange(1, gx * gy + 1), elec)
            elec = np.concatenate((elec, np.random.choice(rest, nElec - elec.shape[0],
                                                          replace=False)))
    _, sig_idx, _ = np.intersect1d(sigChan, elec, return_indices=True)
    if len(sig_idx) == 0:        # the reference retries with the electrode count as "pitch"
        return pitch_sig_channels(chanMap, sigChan, nElec, mmX, mmY, maxElec)
    return sig_idx


def pitch_subsample_sig_channels(pt, pitch, data_path):
    import scipy.io as sio
    chanMap = sio.loadmat(f'{data_path}/{pt}/{pt}_channelMap.mat')['chanMap']
    sigChan = sio.loadmat(f'{data_path}/{pt}/{pt}_sigChannel.mat')['sigChannel']
    for names, (mmX, mmY, maxElec) in _ARRAYS.items():
        if pt in names:
            return pitch_sig_channels(chanMap, sigChan, pitch, mmX, mmY, maxElec)
    raise KeyError('unknown array geometry for subject %s' % pt)
