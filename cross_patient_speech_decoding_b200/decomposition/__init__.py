"""Dimensionality-reduction wrappers, mirroring ``aligned_decoding.decomposition``."""
