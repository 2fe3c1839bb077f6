"""The strict comparison helper itself (tests/parity.py), on CPU: flagged bins are taken out exactly,
anything else is held to the bar."""
import numpy as np
import pytest

from oracle import dsp, nets, pipeline
from parity import strict_check, istft_rows


def _case():
    N, hop = 512, 256
    params = nets.make_synthetic_params("dsd", N // 2 + 1, seed=3)
    mix, _ = pipeline.synth_mixture(1.5, 12)
    want, mag, ph, mm = pipeline.separate(mix, params, "dsd", frameSize=N, hopSize=hop, overlap=25, count_kinks=True,
                                          return_spec=True)
    T = ph.shape[0]
    S = (mm[:, :T] / 0.3) * np.sqrt(N) * np.exp(1j * ph)[None]
    return N, hop, mix, want, mag, ph, mm, S


def test_istft_rows_is_the_oracle_istft_on_sparse_input():
    N, hop, mix, want, mag, ph, mm, S = _case()
    D = np.zeros_like(S[0])
    rows = [3, 4, 17]
    D[rows] = S[0][rows]
    full = dsp.istft_norm(D, window=np.hanning(N), analysisWindow=np.hanning(N), hopsize=float(hop), nfft=float(N))
    np.testing.assert_allclose(istft_rows(D, rows, np.hanning(N), hop, N, mix.size), full[:mix.size], rtol=0, atol=1e-15)


def test_flagged_bins_are_excluded_and_everything_else_is_not():
    N, hop, mix, want, mag, ph, mm, S = _case()
    kmap = np.zeros(ph.shape, dtype=bool)
    X = (mag / 0.3) * np.sqrt(N)
    # a "device" that agrees with the oracle except at two bins where it lands on the other side of the mask jump
    S_dev = S.copy()
    bins = [(20, 40), (33, 7)]
    for t, f in bins:
        S_dev[:, t, f] = 0.25 * X[t, f] * np.exp(1j * ph[t, f])
        kmap[t, f] = True
    got = np.stack([dsp.istft_norm(S_dev[s], window=np.hanning(N), analysisWindow=np.hanning(N), hopsize=float(hop),
                                   nfft=float(N))[:mix.size] for s in range(4)]).astype(np.float32)
    errs = strict_check(None, got, S_dev.astype(np.complex64), want, mag, ph, mm, kmap, N, hop, np.hanning, 0.3)
    assert max(errs) < 2e-7                       # only the float32 rounding of `got` is left
    # the same deviation at a bin the oracle does NOT flag must fail the comparison
    kmap2 = kmap.copy()
    kmap2[20, 40] = False
    if np.linalg.norm(got[0].astype(np.float64) - want[0]) / np.linalg.norm(want[0]) > 1e-4:
        with pytest.raises(AssertionError):
            strict_check(None, got, S_dev.astype(np.complex64), want, mag, ph, mm, kmap2, N, hop, np.hanning, 0.3)
    # an inadmissible value at a flagged bin (mask > 1) is rejected
    S_bad = S_dev.copy()
    S_bad[1, 33, 7] = 1.5 * X[33, 7]
    with pytest.raises(AssertionError):
        strict_check(None, got, S_bad.astype(np.complex64), want, mag, ph, mm, kmap, N, hop, np.hanning, 0.3)
