"""Host logic of deepconvsep_b200.evaluation (Gram assembly from correlation lags, solves, energy
identities, ordering) against the explicit-decomposition restatement in oracle.bsseval, with the
lag table computed by numpy instead of libdcs -- no GPU needed."""
import numpy as np

from oracle import bsseval
from deepconvsep_b200 import evaluation


def numpy_lags(a, b, flen):
    """out[li] = sum_t a[t + li - (flen-1)] * b[t]  (the contract of dcs_xcorr_lags, include/dcs.h)"""
    L = len(a)
    out = np.empty(2 * flen - 1)
    for li in range(2 * flen - 1):
        m = li - (flen - 1)
        out[li] = np.dot(a[m:], b[:L - m]) if m >= 0 else np.dot(a[:L + m], b[-m:])
    return out


def test_energy_identities_match_the_explicit_decomposition():
    rng = np.random.default_rng(11)
    n, L, flen = 3, 5000, 24
    s = np.array([np.convolve(rng.standard_normal(L + 20), rng.standard_normal(5 + 2 * k))[10:10 + L] for k in range(n)])
    est = (s + 0.3 * rng.standard_normal(s.shape) + 0.2 * s[::-1])[[1, 2, 0]]
    s32, e32 = s.astype(np.float32), est.astype(np.float32)
    kinds, idx = evaluation.pair_list(n)
    sig = {"ss": (s32, s32), "se": (s32, e32), "ee": (e32, e32)}
    R = np.array([numpy_lags(sig[k][0][i].astype(np.float64), sig[k][1][j].astype(np.float64), flen) for k, i, j in kinds])
    got = evaluation.ratios_from_lags(R, idx, n, flen)
    want = bsseval.bss_eval_sources(e32.astype(np.float64), s32.astype(np.float64), flen=flen)
    assert list(got[3]) == list(want[3]) == [2, 0, 1]
    for g, w in zip(got[:3], want[:3]):
        assert np.max(np.abs(g - w)) < 1e-8, (g, w)


def test_images_variant_matches_the_explicit_decomposition():
    """multichannel bss_eval_images (DSD100_eval_only.m:240-306): identities on lags vs explicit projections"""
    rng = np.random.default_rng(5)
    nsrc, nchan, L, flen = 3, 2, 2500, 12
    S = np.array([[np.convolve(rng.standard_normal(L + 20), rng.standard_normal(4 + j + c))[10:10 + L]
                   for c in range(nchan)] for j in range(nsrc)]).astype(np.float32)
    E = (S + 0.3 * rng.standard_normal(S.shape) + 0.2 * S[::-1]).astype(np.float32)
    kinds, idx = evaluation.image_pair_list(nsrc, nchan)
    sig = {"ss": (S, S), "se": (S, E), "ee": (E, E)}
    R = np.array([numpy_lags(sig[k][0][a // nchan, a % nchan].astype(np.float64),
                             sig[k][1][b // nchan, b % nchan].astype(np.float64), flen) for k, a, b in kinds])
    got = evaluation.images_from_lags(R, idx, nsrc, nchan, flen)
    want = bsseval.bss_eval_images(E.astype(np.float64), S.astype(np.float64), flen=flen)
    for name, g, w in zip(("SDR", "ISR", "SIR", "SAR"), got, want):
        assert np.max(np.abs(g - w)) < 1e-8, (name, g, w)
    assert evaluation.window_starts(1000, 300, 150) == bsseval.window_starts(1000, 300, 150) == [0, 150, 300, 450, 600]
    assert evaluation.window_starts(100, 300, 150) == []
