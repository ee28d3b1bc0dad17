"""Cross-patient pooling decoders, mirroring ``aligned_decoding.decoders``."""
