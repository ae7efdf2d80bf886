"""Tensor-core 'valid' convolution (csrc/conv_tc.cu: banded Toeplitz GEMM, 3xTF32 on tcgen05) against the CPU oracle
and against the direct shared-memory kernel, at the float tolerance of north_star (1e-5 of the waveform scale)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(x, k, tc, mode="v"):
    import torch

    import dspeed_b200.processors as P

    old = P.TC_CONV_MIN_TAPS
    P.TC_CONV_MIN_TAPS = 1 if tc else 0
    try:
        xd = torch.from_numpy(x).cuda()
        p = {"v": x.shape[1] - len(k) + 1, "f": x.shape[1] + len(k) - 1, "s": x.shape[1]}[mode]
        out = torch.empty((x.shape[0], p), dtype=torch.float32, device="cuda")
        P.convolve_wf(xd, torch.from_numpy(k).cuda(), np.int8(ord(mode)), out)
        torch.cuda.synchronize()
        return out.cpu().numpy()
    finally:
        P.TC_CONV_MIN_TAPS = old


@pytest.mark.parametrize("rows,L,K", [(256, 8192, 256), (200, 8192, 1000), (128, 8192, 4096), (300, 2048, 133), (129, 1024, 1024)])
def test_tc_convolution_matches_oracle(rows, L, K):
    from oracle import oracle as O

    rng = np.random.default_rng(K + rows)
    x = (rng.normal(0, 4, (rows, L)) + 2000 * (np.arange(L)[None, :] > rng.integers(L // 4, 3 * L // 4, (rows, 1)))).astype(np.float32)
    k = rng.standard_normal(K).astype(np.float32)
    ref = O.convolve_wf(x, k, "v")
    got = _run(x, k, tc=True)
    direct = _run(x, k, tc=False)
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 1e-5 * scale, np.abs(got - ref).max() / scale
    assert np.abs(direct - ref).max() <= 1e-5 * scale
    # 3xTF32 is as accurate as the float32 direct kernel to within a small factor
    # measured: 3xTF32 with windowed accumulation stays within 3e-6 of the scale (direct float32 kernel: 5e-7)
    assert np.abs(got - ref).max() <= 4e-6 * scale, np.abs(got - ref).max() / scale


def test_nan_rows_and_structured_kernel():
    from oracle import oracle as O

    rng = np.random.default_rng(3)
    x = rng.normal(0, 10, (256, 4096)).astype(np.float32)
    x[5, 100] = np.nan
    x[77, 4000] = np.nan
    k = np.exp(-np.arange(700) / 150.0).astype(np.float32)
    ref = O.convolve_wf(x, k, "v")
    got = _run(x, k, tc=True)
    assert np.isnan(got[5]).all() and np.isnan(got[77]).all()
    ok = ~np.isnan(ref).any(axis=1)
    assert ok.sum() == 254
    assert np.abs(got[ok] - ref[ok]).max() <= 1e-5 * np.abs(ref[ok]).max()


@pytest.mark.parametrize("mode,rows,L,K", [("f", 150, 4096, 300), ("s", 150, 4096, 301), ("s", 130, 8192, 134), ("f", 128, 1024, 1024)])
def test_full_and_same_modes(mode, rows, L, K):
    """the zero padding of numpy.convolve's 'full' / 'same' is TMA's out-of-bounds fill"""
    from oracle import oracle as O

    rng = np.random.default_rng(K)
    x = (rng.normal(0, 4, (rows, L)) + 1500 * (np.arange(L)[None, :] > rng.integers(L // 4, 3 * L // 4, (rows, 1)))).astype(np.float32)
    k = rng.standard_normal(K).astype(np.float32)
    ref = O.convolve_wf(x, k, mode)
    got = _run(x, k, tc=True, mode=mode)
    assert got.shape == ref.shape
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 4e-6 * scale, np.abs(got - ref).max() / scale
    assert np.abs(_run(x, k, tc=False, mode=mode) - ref).max() <= 1e-5 * scale


def test_pedestal_under_a_zero_area_kernel():
    """A waveform pedestal under a kernel of (almost) zero area: the partial sums of the GEMM are far larger than the
    result, and the tensor core truncates when it accumulates.  The kernel removes the pedestal c[r] = mean(x[r, 0:64]) before
    the products and adds c * sum(kern) back (exact), so the error stays at the float32 level of the RESULT."""
    from oracle import oracle as O

    rng = np.random.default_rng(11)
    x = (rng.normal(0, 4, (256, 8192)) + 3000.0).astype(np.float32)
    k = rng.standard_normal(512).astype(np.float32)
    k -= k.mean()
    ref = np.stack([np.convolve(r.astype(np.float64), k.astype(np.float64), "valid") for r in x[:64]])
    got = _run(x, k, tc=True)[:64]
    # scale of the problem: |x| * ||k||_1 (the result itself is ~1000 x smaller)
    assert np.abs(got - ref).max() <= 2e-7 * 3000.0 * np.abs(k).sum()
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()
