"""Presence score of each propagated object along z (behaviour of REF saber/filters/estimate_thickness.py:
7-112, pinned by the reference-run golden ``tests/golden/saber3d_fit_boundaries.npz``).

Specification (per object column of the ``[Z, nMasks]`` object-score table):
  1. rectify: ``s = max(score, 0)``; baseline = mean of frames ``[-15:-5]``; ``s = max(s - baseline, 0)``;
  2. fit two bump models to ``s(z)`` by bounded non-linear least squares (scipy ``curve_fit`` -> trust-region reflective):
       clipped parabola  d * max(a (z - b)^2 + c, 0),  start (-1e-3, argmax s[1:-1], 1, max(s) / 2),
                         a <= 0, 0 <= b <= Z, 0 <= c <= 10, 0 <= d <= 10
       Gaussian          a * exp(-(z - b)^2 / (2 c^2)),  start (max(s), argmax s[1:-1], 0.3),
                         a >= 0, 0 <= b <= Z, 0 <= c <= 0.25 Z / 2.355   (FWHM at most a quarter of the stack)
  3. a model whose fit raises scores R^2 = 0; the model with the larger R^2 is evaluated on the frame grid (ties and a
     failed parabola go to the Gaussian); both failing gives zeros.
The table is a few kB of float64: host logic (scipy), not a GPU kernel.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Sequence, Tuple

import numpy as np
from scipy.optimize import curve_fit


def _parabola(z, a, b, c, d):
    return d * np.maximum(a * (z - b) ** 2 + c, 0)


def _bell(z, a, b, c):
    with np.errstate(over="ignore"):
        return a * np.exp(-(z - b) ** 2 / (2 * c ** 2))


@dataclass(frozen=True)
class _BumpModel:
    name: str
    fn: Callable
    start: Callable[[np.ndarray, int], Sequence[float]]          # (profile, peak index) -> initial parameters
    bounds: Callable[[int], Tuple[Sequence[float], Sequence[float]]]  # Z -> (lower, upper)


_MODELS = (
    _BumpModel("Quadratic", _parabola,
               lambda s, peak: [-1e-3, peak, 1, np.max(s) / 2],
               lambda Z: ([-np.inf, 0, 0, 0], [0, Z, 10, 10])),
    _BumpModel("Gaussian", _bell,
               lambda s, peak: [np.max(s), peak, 3e-1],
               lambda Z: ((0, 0, 0), (np.inf, Z, Z * 0.25 / 2.355))),
)


def _r2(profile: np.ndarray, fitted: np.ndarray) -> float:
    total = np.sum((profile - np.mean(profile)) ** 2)
    return 0 if total == 0 else 1 - np.sum((profile - fitted) ** 2) / total


def rectify(scores: np.ndarray) -> np.ndarray:
    s = np.maximum(scores, 0)
    s = s - np.mean(s[-15:-5])
    return np.maximum(s, 0)


def _fit(model: _BumpModel, z: np.ndarray, profile: np.ndarray, column: int):
    """-> (R^2, fitted curve or None). The frame grid is float32, as the reference builds it."""
    try:
        params, _ = curve_fit(model.fn, z, profile, p0=model.start(profile, np.argmax(profile[1:-1])),
                              bounds=model.bounds(profile.shape[0]))
    except Exception as e:  # the reference reports and carries on with R^2 = 0
        print(f"Error fitting {model.name} mask {column}: {e}")
        return 0, None
    return _r2(profile, model.fn(np.arange(len(profile)), *params)), model.fn(z, *params)


def fit_organelle_boundaries(frame_scores: np.ndarray, plot: bool = False) -> np.ndarray:
    Z, n = frame_scores.shape
    out = np.zeros((Z, n))
    for col in range(n):
        profile = rectify(frame_scores[:, col].astype(np.float64, copy=True))
        z = np.arange(Z, dtype=np.float32)
        (r2_par, par), (r2_bell, bell) = (_fit(m, z, profile, col) for m in _MODELS)
        if r2_par == 0 and r2_bell == 0:
            continue
        out[:, col] = par if r2_par > r2_bell else bell
    return out
