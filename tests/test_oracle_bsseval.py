"""oracle.bsseval (BSS-Eval 3.0 bss_eval_sources restated) pinned by properties -- there is no MATLAB /
Octave here to run evaluation/bss_eval/bss_eval_sources.m itself (the oracle header says so)."""
import numpy as np
import pytest

from oracle import bsseval


def _sources(n, L, seed):
    rng = np.random.default_rng(seed)
    # coloured, mutually independent sources (white noise through different short filters)
    out = []
    for k in range(n):
        w = rng.standard_normal(L + 64)
        h = rng.standard_normal(8 + 3 * k)
        out.append(np.convolve(w, h, mode="full")[32:32 + L])
    return np.array(out)


def test_project_matches_time_domain_least_squares():
    """`project` (FFT Gram matrix + solve, bss_eval_sources.m:110-159) against the explicit delayed-copy
    design matrix and numpy lstsq."""
    rng = np.random.default_rng(0)
    n, L, flen = 3, 300, 8
    S = rng.standard_normal((n, L))
    se = rng.standard_normal(L)
    m = L + flen - 1
    cols = []
    for k in range(n):
        for a in range(flen):
            c = np.zeros(m)
            c[a:a + L] = S[k]
            cols.append(c)
    A = np.array(cols).T
    y = np.concatenate([se, np.zeros(flen - 1)])
    want = A @ np.linalg.lstsq(A, y, rcond=None)[0]
    got = bsseval.project(se, S, flen)
    assert got.shape == (m,)
    assert np.linalg.norm(got - want) <= 1e-9 * np.linalg.norm(want)


def test_filtered_true_source_is_not_distortion():
    """an estimate that is the true source through a short FIR filter has no interference and no
    artifacts: the decomposition allows a 512-tap time-invariant distortion (bss_eval_sources.m:7-8)."""
    s = _sources(2, 6000, 1)
    s[:, -8:] = 0.0                       # so that truncating the convolution to L samples loses nothing
    h = np.array([0.9, 0.0, -0.3, 0.1])
    se = np.array([np.convolve(s[0], h)[:6000], np.convolve(s[1], h[::-1])[:6000]])
    sdr, sir, sar, perm = bsseval.bss_eval_sources(se, s, flen=64)
    assert list(perm) == [0, 1]
    assert sdr.min() > 100 and sir.min() > 100 and sar.min() > 100


def test_known_leak_gives_the_analytic_sir():
    s = _sources(2, 20000, 2)
    s /= np.sqrt(np.mean(s ** 2, axis=1, keepdims=True))
    alpha = 0.1
    se = np.array([s[0] + alpha * s[1], s[1]])
    sdr, sir, sar, perm = bsseval.bss_eval_sources(se, s, flen=32)
    assert abs(sir[0] - 20.0) < 0.5          # -20 log10(alpha), up to the finite-length cross-correlation
    assert sar[0] > 100                       # se lies in the span of the sources: no artifacts
    assert abs(sdr[0] - sir[0]) < 1e-6        # so SDR == SIR
    assert sir[1] > 100 and sdr[1] > 100      # the untouched source


def test_permutation_and_gain_invariance():
    s = _sources(3, 8000, 3)
    rng = np.random.default_rng(4)
    est = s + 0.2 * rng.standard_normal(s.shape) * s.std(axis=1, keepdims=True)
    base = bsseval.bss_eval_sources(est, s, flen=32)
    order = [2, 0, 1]
    sw = bsseval.bss_eval_sources(est[order], s, flen=32)
    # estimate perm[j] matches true source j
    assert [order[p] for p in sw[3]] == [0, 1, 2]
    for a, b in zip(base[:3], sw[:3]):
        assert np.allclose(a, b, atol=1e-9)
    # a gain on an estimate is a (1-tap) filter distortion: the ratios do not move
    sc = bsseval.bss_eval_sources(est * np.array([[0.5], [2.0], [-1.0]]), s, flen=32)
    for a, b in zip(base[:3], sc[:3]):
        assert np.allclose(a, b, atol=1e-8)
    # additive noise of -14 dB: SDR in the expected range
    assert 10 < base[0].min() < base[0].max() < 18


def test_shape_errors():
    with pytest.raises(ValueError):
        bsseval.bss_eval_sources(np.zeros((2, 10)), np.zeros((3, 10)))


def test_images_variant_projection_and_single_channel_consistency():
    """multichannel `project` of DSD100_eval_only.m:257-295 against brute-force least squares; with one
    channel SIR and SAR of the images variant are those of bss_eval_sources for the matching pair"""
    rng = np.random.default_rng(8)
    nsrc, nchan, L, flen = 2, 2, 400, 6
    S = rng.standard_normal((nsrc, nchan, L))
    se = rng.standard_normal((nchan, L))
    m = L + flen - 1
    cols = []
    for j in range(nsrc):
        for c in range(nchan):
            for a in range(flen):
                v = np.zeros(m)
                v[a:a + L] = S[j, c]
                cols.append(v)
    A = np.array(cols).T
    got = bsseval.project_images(se, S, flen)
    for ch in range(nchan):
        y = np.concatenate([se[ch], np.zeros(flen - 1)])
        want = A @ np.linalg.lstsq(A, y, rcond=None)[0]
        assert np.linalg.norm(got[ch] - want) <= 1e-9 * np.linalg.norm(want)
    s = _sources(3, 6000, 9)
    est = s + 0.2 * rng.standard_normal(s.shape) * s.std(axis=1, keepdims=True) + 0.1 * s[::-1]
    sdr_i, isr_i, sir_i, sar_i = bsseval.bss_eval_images(est[:, None, :], s[:, None, :], flen=32)
    sdr_s, sir_s, sar_s, perm = bsseval.bss_eval_sources(est, s, flen=32)
    assert list(perm) == [0, 1, 2]
    assert np.allclose(sir_i, sir_s, atol=1e-9) and np.allclose(sar_i, sar_s, atol=1e-9)
    # SDR differs by definition (images: true source vs everything else; sources: filtered source)
    assert np.all(sdr_i <= sdr_s + 1e-9)
    # a perfect estimate: no distortion of any kind
    r = bsseval.bss_eval_images(s[:, None, :], s[:, None, :], flen=8)
    assert min(x.min() for x in r) > 100
    # windows
    w = bsseval.bss_eval_windowed(est[:, None, :], s[:, None, :], 3000, 1500, flen=8)
    assert w[0].shape == (3, 3)
    assert np.allclose(w[0][:, 1], bsseval.bss_eval_images(est[:, None, 1500:4500], s[:, None, 1500:4500], flen=8)[0])
