"""Electrode-subsampling front-end (drop-in for the reference's ``processing_utils``
index generators used by scripts/aligned_decode_{grid,pitch,spatialAvg}_subsample.py).
Index generation is host integer code that reproduces the reference's results (and its
numpy RNG call order) bit for bit; the channel gather / block averaging of the resident
trials runs on the GPU (``device_subsample``)."""
