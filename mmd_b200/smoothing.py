"""Savitzky-Golay smoothing of trajectory batches on the device.

Reference: mmd/common/trajectory_utils.py:31-38 -> scipy.signal.savgol_filter(x, window_length=10, polyorder=2, axis=1), i.e.
mode='interp'.  scipy (third-party, unpinned in the reference's requirements.txt; 1.18 here) computes
  * interior samples: a correlation with the least-squares coefficients of a degree-`polyorder` fit over `window_length`
    samples evaluated at pos = window_length // 2 - 0.5 for an even window (scipy/signal/_savitzky_golay.py, savgol_coeffs),
  * the first / last window_length // 2 samples: a polynomial fitted to the first / last `window_length` samples and
    evaluated at those positions (_fit_edges_polyfit).
Every output sample is a fixed linear combination of input samples, so the whole filter is ONE matrix S [H, H]; it is
built here in float64 with numpy (no scipy at run time) and applied by mmdk_smooth_trajs (double accumulation).
tests/test_host_logic.py pins S against scipy itself.
"""
import ctypes as C
import functools

import numpy as np
import torch

from . import _lib


def savgol_coeffs(window_length, polyorder):
    """scipy.signal.savgol_coeffs(window_length, polyorder) (deriv=0, use='conv')."""
    halflen, rem = divmod(window_length, 2)
    pos = halflen - 0.5 if rem == 0 else float(halflen)
    x = np.arange(-pos, window_length - pos, dtype=np.float64)[::-1]
    A = x ** np.arange(polyorder + 1, dtype=np.float64).reshape(-1, 1)
    y = np.zeros(polyorder + 1)
    y[0] = 1.0
    coeffs, *_ = np.linalg.lstsq(A, y, rcond=None)
    return coeffs


@functools.lru_cache(maxsize=16)
def savgol_matrix(n, window_length=10, polyorder=2):
    """S [n, n] float64 with savgol_filter(x, window_length, polyorder, mode='interp') == S @ x."""
    if window_length > n:
        raise ValueError("window_length must be <= the number of samples (scipy raises too)")
    coeffs = savgol_coeffs(window_length, polyorder)
    S = np.zeros((n, n))
    # interior: scipy.ndimage.convolve1d(x, coeffs, mode='constant'): y[i] = sum_j coeffs[j] x[i + w // 2 - j], so for an even
    # length the window of output i covers inputs i - (w/2 - 1) .. i + w/2
    w = window_length
    rev = coeffs[::-1]
    lo = w - 1 - w // 2
    for i in range(n):
        for j in range(w):
            src = i - lo + j
            if 0 <= src < n:
                S[i, src] += rev[j]
    # edges: least-squares polynomial over the first / last w samples, evaluated at the first / last w // 2 positions
    halflen = w // 2
    t = np.arange(w, dtype=np.float64)
    V = np.vander(t, polyorder + 1, increasing=True)          # [w, p+1]
    P = np.linalg.pinv(V)                                      # poly coefficients = P @ x_window
    for i in range(halflen):
        S[i, :] = 0.0
        S[i, :w] = (np.vander(np.array([float(i)]), polyorder + 1, increasing=True) @ P)[0]
        k = n - halflen + i
        S[k, :] = 0.0
        S[k, n - w:] = (np.vander(np.array([float(w - halflen + i)]), polyorder + 1, increasing=True) @ P)[0]
    return S


_dev_cache = {}


def smooth_trajs(trajs, window_size=10, poly_order=2):
    """mmd/common/trajectory_utils.py:31-38 for a [B, H, D] CUDA batch (or a list of [H, D] tensors): stays on the device."""
    if not torch.is_tensor(trajs):
        return [smooth_trajs(t[None], min(window_size, t.shape[0]), poly_order)[0] if min(window_size, t.shape[0]) > 2 else t
                for t in trajs]
    assert trajs.dim() == 3
    lib = _lib.lib()
    x = trajs.to(torch.float32).contiguous()
    B, H, D = x.shape
    key = (H, window_size, poly_order, str(x.device))
    if key not in _dev_cache:
        _dev_cache[key] = torch.from_numpy(savgol_matrix(H, window_size, poly_order)).to(x.device).contiguous()
    out = torch.empty_like(x)
    _lib.check(lib.mmdk_smooth_trajs(_lib.ptr(_dev_cache[key]), _lib.ptr(x), B, H, D, _lib.ptr(out), _lib.stream_ptr()))
    return out
