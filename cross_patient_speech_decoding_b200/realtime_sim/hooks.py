"""reduce_to_latent_space / align_to_target (realtime_sim/realtime_datamodule.py:813-894) on the
package's GPU classes (decomposition.PCA.PCA, alignment.AlignCCA.AlignCCA or any aligner with
the same fit / transform signature)."""
import numpy as np
import torch

from ..decomposition.PCA import PCA


def reduce_to_latent_space(data, pca=None, n_components=30, low_thresh=5):
    """data: (N, T, C) tensor / array -> ((N, T, k) float32 tensor, fitted PCA).  ``pca`` given:
    transform only.  A fit that keeps <= ``low_thresh`` components (a variance threshold
    swallowed by one artifact component) is redone with 30 components, as the reference does
    (realtime_datamodule.py:855-872; note that, like the reference, the re-fit keeps all 30
    components -- the "drop the first" in its comment is not in its code)."""
    arr = data.detach().cpu().numpy() if isinstance(data, torch.Tensor) else np.asarray(data)
    shp = arr.shape
    flat = arr.reshape(-1, shp[-1])
    if pca is not None:
        dr = pca
        red = dr.transform(flat)
    else:
        dr = PCA(n_components=n_components)
        red = dr.fit_transform(flat)
        if dr.n_components_ <= low_thresh:
            dr = PCA(n_components=30)
            red = dr.fit_transform(flat)
    return torch.Tensor(np.asarray(red).reshape(shp[0], shp[1], -1)), dr


def align_to_target(aligner, target_data, source_data, target_labels, source_labels):
    """Fits ``aligner()`` on (target, source) and maps the source trials into the target's
    space (realtime_datamodule.py:875-894)."""
    def host(x):
        return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    src = host(source_data)
    shp = src.shape
    align = aligner()
    align.fit(host(target_data), src, host(target_labels), host(source_labels))
    out = align.transform(src)
    return torch.Tensor(np.asarray(out).reshape(shp[0], shp[1], -1))
