"""TEST INFRASTRUCTURE (checker) -- CPU restatement of the remaining processors of the reference's SiPM chain
(tests/configs/sipm-dsp-config.json) and of the kernel generators that round 1 left out, in numpy / plain Python
(small cases only).  Pinned bit-for-bit (integers, indices, histogram weights) / to float rounding against
tests/golden/sipm_processors.npz, which oracle/gen_golden.py recorded from the reference's own numba processors.
Never imported by the product path.

Type handling follows what numba does with the reference's code: in the float32 loop a float32 / int quotient and
``np.linspace`` are float64, float32 (-) float32 stays float32, results are rounded when stored into float32 outputs.
Every function takes 2-D ``[rows, n]`` (or 1-D) arrays and loops over rows like the gufunc does.
"""
from __future__ import annotations

import numpy as np


def _rows(a):
    a = np.asarray(a)
    return a[None, :] if a.ndim == 1 else a


def gaussian_filter1d(sigma, truncate, dtype=np.float64):
    """gaussian_filter1d.py:46-82"""
    dt = np.dtype(dtype).type
    sigma, truncate = dt(sigma), dt(truncate)
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / sigma2 * x**2)
    phi = phi / phi.sum()
    return np.asarray(phi, dtype=np.float64).astype(dtype)


def reflected_convolve_wf(w_in, kernel):
    """convolutions.py:122-182: 'same' convolution of the reflect-padded waveform, cut back to the input length"""
    w = _rows(w_in)
    kernel = np.asarray(kernel)
    out = np.full(w.shape, np.nan, w.dtype)
    ext = int(len(kernel) / 2) + 1
    for r in range(w.shape[0]):
        if np.isnan(w[r]).any() or np.isnan(kernel).any():
            continue
        e = np.pad(w[r], ext, mode="reflect")
        out[r] = np.convolve(e, kernel, mode="same")[ext:-ext]
    return out


def _linspace(start, stop, num):
    # numba's np.linspace: the end points are taken as float64 (pinned by the float32 goldens)
    arr = np.empty(num, np.float64)
    div = num - 1
    if div > 0:
        delta = np.float64(stop) - np.float64(start)
        step = delta / np.float64(div)
        for i in range(num):
            arr[i] = np.float64(start) + i * step
    else:
        arr[0] = start
    if num > 1:
        arr[-1] = stop
    return arr


def histogram(w_in, n_bins):
    """histogram.py:14-89 -> (weights [rows, n_bins], borders [rows, n_bins + 1])"""
    w = _rows(w_in)
    dt = w.dtype.type
    weights = np.zeros((w.shape[0], n_bins), w.dtype)
    borders = np.full((w.shape[0], n_bins + 1), np.nan, w.dtype)
    for r in range(w.shape[0]):
        x = w[r]
        if np.isnan(x).any():
            continue
        wf_min, wf_max = x.min(), x.max()
        delta = np.float64(wf_max - wf_min) / np.float64(n_bins)
        borders[r] = _linspace(wf_min, wf_max, n_bins + 1).astype(w.dtype)
        if delta == 0:
            continue
        b0 = borders[r, 0]
        for v in x:
            if v == wf_max:
                continue
            k = int(np.floor(np.float64(dt(v - b0)) / delta))
            if 0 <= k < n_bins:
                weights[r, k] += 1
    return weights, borders


def histogram_around_mode(w_in, center, bin_width, n_bins):
    """histogram.py:92-204 (numba unifies `center` to float64; bin_width stays in the loop's type)"""
    w = _rows(w_in)
    dt = w.dtype.type
    bw = dt(bin_width)
    weights = np.zeros((w.shape[0], n_bins), w.dtype)
    borders = np.zeros((w.shape[0], n_bins + 1), w.dtype)
    for r in range(w.shape[0]):
        x = w[r]
        if np.isnan(x).any():
            raise ValueError("input data contains nan")
        c = np.float64(dt(center))
        if np.isnan(c):
            tmp = np.zeros(n_bins, w.dtype)
            wf_min, wf_max = x.min(), x.max()
            delta = np.float64(wf_max - wf_min) / np.float64(n_bins)
            b = _linspace(wf_min, wf_max, n_bins + 1).astype(w.dtype)
            if delta == 0:
                c = np.float64(wf_min)
            else:
                for v in x:
                    if v == wf_max:
                        continue
                    k = int(np.floor(np.float64(dt(v - b[0])) / delta))
                    if 0 <= k < n_bins:
                        tmp[k] += 1
                c = np.float64(b[np.argmax(tmp)]) + 0.5 * delta
                c = np.round(c / np.float64(bw)) * np.float64(bw)
        hist_min = c - np.float64(bw) * (n_bins // 2) - 0.5 * np.float64(bw)
        borders[r] = (hist_min + np.float64(bw) * np.arange(n_bins + 1)).astype(w.dtype)
        b0 = borders[r, 0]
        for v in x:
            k = int(np.floor(dt(dt(v - b0) / bw)))
            if 0 <= k < n_bins:
                weights[r, k] += 1
    return weights, borders


def histogram_stats(weights_in, edges_in, max_in):
    """histogram_stats.py:146-261 -> (mode index, left edge of the mode bin, half width) per row"""
    wts, edg = _rows(weights_in), _rows(edges_in)
    dt = wts.dtype.type
    n_rows, nb = wts.shape
    mode = np.full(n_rows, np.nan, wts.dtype)
    mx = np.full(n_rows, np.nan, wts.dtype)
    fw = np.full(n_rows, np.nan, wts.dtype)
    max_in = np.broadcast_to(np.asarray(max_in, wts.dtype), (n_rows,))
    for r in range(n_rows):
        w, e = wts[r], edg[r]
        if np.isnan(w).any():
            continue
        mi = 0
        if np.isnan(max_in[r]):
            for i in range(nb):
                if w[i] > w[mi]:
                    mi = i
        elif max_in[r] > e[-2]:
            mi = nb - 1
        else:
            for i in range(nb):
                if abs(max_in[r] - e[i]) < abs(max_in[r] - e[mi]):
                    mi = i
        mode[r] = mi
        mx[r] = e[mi]
        for i in range(mi, nb):
            if w[i] <= 0.5 * w[mi] and w[i] != 0:
                fw[r] = abs(mx[r] - e[i])
                break
        for i in range(0, mi):
            if w[i] >= 0.5 * w[mi] and w[i] != 0:
                if fw[r] < abs(mx[r] - e[i]):
                    fw[r] = abs(mx[r] - e[i])
                break
    return mode, mx, fw


def histogram_peakstats(weights_in, edges_in, max_in, skip_zeroes, width_type):
    """histogram_stats.py:12-143 -> (mode = centre of the mode bin, width) per row"""
    wts, edg = _rows(weights_in), _rows(edges_in)
    dt = wts.dtype.type
    n_rows, nb = wts.shape
    mode = np.full(n_rows, np.nan, wts.dtype)
    width = np.full(n_rows, np.nan, wts.dtype)
    max_in = np.broadcast_to(np.asarray(max_in, wts.dtype), (n_rows,))
    half = np.float64(0.5)
    for r in range(n_rows):
        w, e = wts[r], edg[r]
        mi = 0
        if np.isnan(max_in[r]):
            for i in range(nb):
                if w[i] > w[mi]:
                    mi = i
        elif max_in[r] > e[-1]:
            mi = nb - 1
        elif max_in[r] < e[0]:
            mi = 0
        else:
            for i in range(nb):
                if e[i] <= max_in[r] < e[i + 1]:
                    mi = i
                    break
        mode[r] = dt(np.float64(e[mi]) + half * np.float64(dt(e[mi + 1] - e[mi])))
        left = right = np.nan
        for i in range(mi, nb):
            if skip_zeroes and w[i] == 0:
                continue
            if w[i] <= half * np.float64(w[mi]):
                right = abs(dt(mode[r] - e[i]))
                break
        else:
            right = abs(dt(mode[r] - e[-1]))
        for i in range(mi, -1, -1):
            if skip_zeroes and w[i] == 0:
                continue
            if w[i] <= half * np.float64(w[mi]):
                left = abs(dt(mode[r] - e[i + 1]))
                break
        else:
            left = abs(dt(mode[r] - e[0]))
        width[r] = {0: np.float64(left) + np.float64(right), 1: min(left, right), 2: max(left, right), 3: left, 4: right}[int(width_type)]
    return mode, width


def peak_snr_threshold(w_in, idx_in, ratio_in, width_in):
    """peak_snr_threshold.py:11-71 -> (idx_out [rows, m] NaN padded, n_idx_out uint32)"""
    w, idx = _rows(w_in), _rows(idx_in)
    dt = w.dtype.type
    n_rows, m = idx.shape
    out = np.full((n_rows, m), np.nan, w.dtype)
    cnt = np.zeros(n_rows, np.uint32)
    n = w.shape[1]
    for r in range(n_rows):
        k = 0
        for i in range(m):
            if not np.isnan(idx[r, i]):
                a = int(idx[r, i]) - int(width_in)
                b = int(idx[r, i]) + int(width_in)
                a = max(a, 0)
                if b >= n:
                    b = n - 1
                mi = a
                for j in range(a, b):
                    if w[r, j] < w[r, mi]:
                        mi = j
                if np.absolute(dt(w[r, mi] / w[r, int(idx[r, i])])) < dt(ratio_in):
                    out[r, k] = idx[r, i]
                    k += 1
        cnt[r] = k
    return out, cnt


def multi_a_filter(w_in, vt_maxs_in):
    """multi_a_filter.py:11-57: the waveform's value at every (integer) time of the list, NaN padded"""
    w, vt = _rows(w_in), _rows(vt_maxs_in)
    n_rows, m = vt.shape
    out = np.full((n_rows, m), np.nan, w.dtype)
    n = w.shape[1]
    for r in range(n_rows):
        if np.isnan(w[r]).any():
            continue
        nan_mask = np.isnan(vt[r])
        if nan_mask.all() or m == 0:
            continue
        first = np.where(nan_mask)[0]
        first = None if len(first) == 0 else first[0]
        if first is not None and (~np.isnan(vt[r, first:])).any():
            first = None
        for i in range(m if first is None else first):
            t = vt[r, i]
            if np.isnan(t) or t < 0 or t >= n:
                continue        # fixed_time_pickoff: NaN outside the waveform (fixed_time_pickoff.py:74-82)
            if np.floor(t) != t:
                raise ValueError("fixed_time_pickoff requires integer t_in when using mode 'i'")
            out[r, i] = w[r, int(t)]
    return out


def dplms(noise_mat, reference, a1, a2, a3, ff, length, dtype=np.float64):
    """energy_kernels.py:160-272 (float64 linear algebra, result cast to the kernel's dtype before the
    normalisation, like the reference's in-place `kernel[:] = ...; kernel[:] /= maxy`)"""
    noise_mat = np.array(noise_mat)
    reference = np.array(reference)
    ssize = len(reference)
    flo = int(ssize / 2 - length / 2)
    fhi = int(ssize / 2 + length / 2)
    ref_mat = np.zeros([length, length])
    ref_sig = np.zeros([length])
    shifts = [0] if ff == 0 else [-1, 0, 1]
    for i in shifts:
        ref_mat += np.outer(reference[flo + i: fhi + i], reference[flo + i: fhi + i])
        ref_sig += reference[flo + i: fhi + i]
    ref_mat /= len(shifts)
    mat = a1 * noise_mat + a2 * ref_mat + a3 * np.ones([length, length])
    kernel = np.zeros(length, dtype)
    kernel[:] = np.flip(np.linalg.solve(mat, ref_sig))
    y = np.convolve(reference, kernel, mode="valid")
    kernel[:] /= np.amax(y)
    return kernel
