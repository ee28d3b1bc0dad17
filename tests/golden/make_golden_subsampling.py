"""Generates tests/golden/subsampling.npz by running the UNMODIFIED reference index
generators (processing_utils/{grid_subsampling,poisson_disk_sampling,spatial_avg_subsampling}.py
imported from /root/reference) on a synthetic 8 x 16 channel map written to temporary .mat
files:    python tests/golden/make_golden_subsampling.py
"""
import os
import sys
import tempfile
import types

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
try:
    import matplotlib  # noqa: F401
except ImportError:            # the reference imports pyplot at module level only for its demos
    mpl = types.ModuleType('matplotlib')
    mpl.pyplot = types.ModuleType('matplotlib.pyplot')
    sys.modules['matplotlib'] = mpl
    sys.modules['matplotlib.pyplot'] = mpl.pyplot
sys.path.insert(0, '/root/reference/aligned_decoding')
from processing_utils import grid_subsampling as ref_grid  # noqa: E402
from processing_utils import poisson_disk_sampling as ref_pds  # noqa: E402
from processing_utils import spatial_avg_subsampling as ref_avg  # noqa: E402


def synthetic_maps():
    """8 x 16 map with channel numbers 1..128 (row-major), two dead corners, and a 24-wide
    variant with NaN borders; 70 'significant' channels."""
    chan = np.arange(1, 129, dtype=float).reshape(8, 16)
    chan[0, 0] = np.nan
    chan[7, 15] = np.nan
    wide = np.full((8, 18), np.nan)        # not 24 wide: stays untrimmed
    wide[:, 1:-1] = chan
    sig = np.sort(np.random.default_rng(3).choice(np.arange(1, 129), 70, replace=False))
    return chan, sig


def main():
    chan, sig = synthetic_maps()
    out = dict(chanMap=chan, sigChan=sig)
    with tempfile.TemporaryDirectory() as root:
        os.makedirs(os.path.join(root, 'S14'))
        sio.savemat(os.path.join(root, 'S14', 'S14_channelMap.mat'), {'chanMap': chan})
        sio.savemat(os.path.join(root, 'S14', 'S14_sigChannel.mat'), {'sigChannel': sig[None, :]})
        for tag, win, step in (('a', (4, 8), (1, 1)), ('b', (3, 5), (2, 3)), ('c', (8, 16), (1, 1))):
            idxs = ref_grid.grid_susbsample_idxs((8, 16), win, step=step)
            out['grid_%s' % tag] = np.stack(idxs)
            lst = ref_grid.grid_subsample_sig_channels('S14', win, root, step=step)
            out['gridsig_%s_n' % tag] = len(lst)
            for i, a in enumerate(lst):
                out['gridsig_%s_%d' % (tag, i)] = a
        for n_elec in (10, 30, 64, 100):
            np.random.seed(100 + n_elec)
            spacing = np.floor(np.sqrt(8 * 16 / n_elec))
            out['pds_%d' % n_elec] = ref_pds.poisson_disk_sampling((8, 16), spacing, n_elec)
        for pitch in (1.0, 1.5, 2.0, 3.0, 5.0):
            np.random.seed(int(pitch * 10))
            key = 'pitch_%s' % str(pitch).replace('.', 'p')
            out[key] = ref_pds.pitch_subsample_sig_channels('S14', pitch, root)
        for cs in (2, 3, 4):
            out['avg_%d' % cs] = np.stack(ref_avg.spatial_avg_idxs((8, 16), cs))
            lst = ref_avg.spatial_avg_sig_channels('S14', cs, root, useSig=True)
            out['avgsig_%d_n' % cs] = len(lst)
            for i, a in enumerate(lst):
                out['avgsig_%d_%d' % (cs, i)] = a
        data = np.random.default_rng(5).standard_normal((6, 8, 16, 25))
        out['avg_data_in'] = data
        out['avg_data_out'] = ref_avg.spatial_avg_data(data, ref_avg.spatial_avg_idxs((8, 16), 3))
    np.savez_compressed(os.path.join(HERE, 'subsampling.npz'), **out)
    print('subsampling.npz:', len(out), 'arrays;', {k: v.shape for k, v in out.items() if k.startswith('pds')})


if __name__ == '__main__':
    main()
