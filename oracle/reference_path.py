"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the reference's own classes from ``$CPSD_REF_PATH`` (default
``/root/reference``) when that tree exists (this container only; the GPU box
has no copy) and provides ``AlignMCCA`` by injecting the restated MCCA
(oracle/mcca_restated.py, PARITY UNPINNED) for the missing ``mvlearn``
dependency -- the reference's wrapper logic (alignment/AlignMCCA.py:13-174)
then runs unmodified on top of it.
"""
import os
import sys
import types

REF_ROOT = os.environ.get('CPSD_REF_PATH', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'aligned_decoding'))


def load():
    """Returns a namespace with the reference's hot-path classes."""
    if not available():
        raise RuntimeError('reference tree not present at %s' % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if 'mvlearn' not in sys.modules:
        try:
            import mvlearn.embed  # noqa: F401
        except Exception:
            from oracle.mcca_restated import MCCARestated
            mv = types.ModuleType('mvlearn')
            emb = types.ModuleType('mvlearn.embed')
            emb.MCCA = MCCARestated
            mv.embed = emb
            sys.modules['mvlearn'] = mv
            sys.modules['mvlearn.embed'] = emb
    ns = types.SimpleNamespace()
    from aligned_decoding.alignment import AlignCCA as m_cca
    from aligned_decoding.alignment import AlignMCCA as m_mcca
    from aligned_decoding.alignment import JointPCA as m_jpca
    from aligned_decoding.alignment import alignment_utils as m_utils
    from aligned_decoding.decomposition import DimRedReshape as m_drr
    from aligned_decoding.decomposition import NoCenterPCA as m_ncp
    from aligned_decoding.decoders import cross_pt_decoders as m_dec
    ns.AlignCCA = m_cca.AlignCCA
    ns.CCA_align = m_cca.CCA_align
    ns.AlignCCA_module = m_cca
    ns.AlignMCCA = m_mcca.AlignMCCA
    ns.AlignMCCA_module = m_mcca
    ns.JointPCA = m_jpca.JointPCA
    ns.utils = m_utils
    ns.DimRedReshape = m_drr.DimRedReshape
    ns.NoCenterPCA = m_ncp.NoCenterPCA
    ns.decoders = m_dec
    return ns
