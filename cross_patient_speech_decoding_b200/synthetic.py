"""Synthetic uECoG-shaped data for tests and benchmarks.

The real recordings used by the reference are not public; this generator makes
patients that share latent per-token trajectories, mixed through per-patient
random read-out matrices, with rank-1 common-mode noise and iid sensor noise.
Shapes follow the reference's data layout ``(trials, time, channels)`` with
per-trial decode labels ``y`` (first phoneme, 1..9) and alignment labels
``y_align`` (3-phoneme sequence), see reference
``alignment/alignment_utils.py:127-184`` for the triple ``(X, y, y_align)``.
"""
import numpy as np

N_TOKENS = 52
N_PHON = 9
N_LATENT = 20
N_BUMPS = 8


def vocabulary():
    """52 three-phoneme tokens, phonemes 1..9 (fixed seed 7)."""
    return np.random.default_rng(7).integers(1, N_PHON + 1, (N_TOKENS, 3))


def shared_latents(n_time=200):
    """Per-token latent trajectories ``(52, T, L)`` (fixed seed 123)."""
    t = np.arange(n_time)[:, None]
    centers = np.linspace(0, n_time - 1, N_BUMPS)[None, :]
    bumps = np.exp(-0.5 * ((t - centers) / (0.12 * n_time)) ** 2)  # (T, 8)
    coef = np.random.default_rng(123).standard_normal(
        (N_TOKENS, N_BUMPS, N_LATENT))
    lat = np.einsum('tb,kbl->ktl', bumps, coef)
    return lat * np.linspace(2.0, 0.3, N_LATENT)[None, None, :]


def make_patient(p, n_trials=144, n_time=200, n_chan=128, noise=0.15,
                 common=0.35, dtype=np.float64):
    """One patient ``(X, y, y_align)``; seed ``1000 + p``."""
    rng = np.random.default_rng(1000 + p)
    vocab = vocabulary()
    lat = shared_latents(n_time)
    tok = rng.integers(0, N_TOKENS, n_trials)
    A = rng.standard_normal((N_LATENT, n_chan)) / np.sqrt(N_LATENT)
    g = rng.standard_normal((n_trials, n_time, 1))
    h = rng.standard_normal((1, 1, n_chan))
    eps = rng.standard_normal((n_trials, n_time, n_chan))
    X = lat[tok] @ A + common * g * h + noise * eps
    y_align = vocab[tok]
    y = y_align[:, 0].copy()
    return np.ascontiguousarray(X, dtype=dtype), y, y_align


def make_patients(n_patients, **kw):
    """List of ``(X, y, y_align)`` for patients ``0..n_patients-1``."""
    return [make_patient(p, **kw) for p in range(n_patients)]
