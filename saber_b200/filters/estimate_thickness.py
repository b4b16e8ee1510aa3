"""Twin of REF saber/filters/estimate_thickness.py:7-112 (R8): per-object presence score along z from the mask-decoder
object-score logits — clamp >= 0, subtract the mean of frames [-15:-5], fit a clipped quadratic and a Gaussian with
scipy's TRF least squares, keep the better R^2. Z x nMasks float64 values (a few kB): host logic, not a GPU kernel."""
from __future__ import annotations

import numpy as np
from scipy.optimize import curve_fit


def quadratic(x, a, b, c, d):
    return d * np.maximum(a * (x - b) ** 2 + c, 0)


def gaussian(x, a, b, c):
    with np.errstate(over="ignore"):
        return a * np.exp(-(x - b) ** 2 / (2 * c ** 2))


def calculate_r2_score(data, func, fit_params):
    x = np.arange(len(data))
    y_fit = func(x, *fit_params)
    ss_res = np.sum((data - y_fit) ** 2)
    ss_tot = np.sum((data - np.mean(data)) ** 2)
    return 0 if ss_tot == 0 else 1 - ss_res / ss_tot


def fit_quadratic(x, data):
    nFrames = data.shape[0]
    x_max = np.argmax(data[1:-1])
    popt, _ = curve_fit(quadratic, x, data, p0=[-1e-3, x_max, 1, np.max(data) / 2],
                        bounds=([-np.inf, 0, 0, 0], [0, nFrames, 10, 10]))
    return popt, calculate_r2_score(data, quadratic, popt)


def fit_gaussian(x, data):
    nFrames = data.shape[0]
    x_max = np.argmax(data[1:-1])
    c_max = nFrames * 0.25 / 2.355
    popt, _ = curve_fit(gaussian, x, data, p0=[np.max(data), x_max, 3e-1], bounds=((0, 0, 0), (np.inf, nFrames, c_max)))
    return popt, calculate_r2_score(data, gaussian, popt)


def preprocess(data: np.ndarray):
    data = np.maximum(data, 0)
    data -= np.mean(data[-15:-5])
    return np.maximum(data, 0)


def fit_organelle_boundaries(frame_scores: np.ndarray, plot: bool = False):
    nFrames, nMasks = frame_scores.shape
    mask_boundaries = np.zeros((nFrames, nMasks))
    for ii in range(nMasks):
        data = preprocess(frame_scores[:, ii].copy())
        x = np.arange(len(data), dtype=np.float32)
        try:
            popt1, r2_quad = fit_quadratic(x, data)
        except Exception as e:  # the reference prints and carries on
            print(f"Error fitting Quadratic mask {ii}: {e}")
            r2_quad = 0
        try:
            popt2, r2_gauss = fit_gaussian(x, data)
        except Exception as e:
            print(f"Error fitting Gaussian mask {ii}: {e}")
            r2_gauss = 0
        if r2_quad == 0 and r2_gauss == 0:
            mask_boundaries[:, ii] = 0
        elif r2_quad > r2_gauss:
            mask_boundaries[:, ii] = quadratic(x, *popt1)
        else:
            mask_boundaries[:, ii] = gaussian(x, *popt2)
    return mask_boundaries
