"""Alignment classes, same module names as the reference's ``aligned_decoding.alignment``."""
