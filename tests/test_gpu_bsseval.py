"""GPU BSS-Eval: the float64 correlation-lag kernel against numpy, and
deepconvsep_b200.evaluation.bss_eval_sources against the oracle restatement of
evaluation/bss_eval/bss_eval_sources.m."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import bsseval  # noqa: E402


def numpy_lags(a, b, flen):
    L = len(a)
    out = np.empty(2 * flen - 1)
    for li in range(2 * flen - 1):
        m = li - (flen - 1)
        if abs(m) >= L:
            out[li] = 0.0
        else:
            out[li] = np.dot(a[m:], b[:L - m]) if m >= 0 else np.dot(a[:L + m], b[-m:])
    return out


@pytest.mark.parametrize("L,flen", [(70001, 512), (65536, 512), (4099, 37), (100, 512), (1, 1)])
def test_xcorr_lags_match_numpy(L, flen):
    from deepconvsep_b200 import evaluation
    from deepconvsep_b200.engine import Context
    rng = np.random.default_rng(L + flen)
    sig = rng.standard_normal((3, L)).astype(np.float32)
    d = torch.tensor(sig, device="cuda")
    pairs = [(0, 0), (1, 0), (2, 1), (0, 2)]
    R = evaluation.xcorr_lags(Context(0), [(d[i], d[j]) for i, j in pairs], L, flen)
    assert R.shape == (4, 2 * flen - 1)
    for q, (i, j) in enumerate(pairs):
        want = numpy_lags(sig[i].astype(np.float64), sig[j].astype(np.float64), flen)
        scale = np.sqrt(np.dot(sig[i].astype(np.float64), sig[i]) * np.dot(sig[j].astype(np.float64), sig[j]))
        assert np.max(np.abs(R[q] - want)) <= 1e-12 * max(scale, 1.0), (i, j)
    # deterministic: a second run gives the same bits
    R2 = evaluation.xcorr_lags(Context(0), [(d[i], d[j]) for i, j in pairs], L, flen)
    assert np.array_equal(R, R2)


@pytest.mark.parametrize("n,L,flen", [(3, 40000, 512), (2, 30011, 64), (4, 50000, 512)])
def test_bss_eval_sources_matches_oracle(n, L, flen):
    from deepconvsep_b200 import evaluation
    rng = np.random.default_rng(n * 1000 + flen)
    s = np.array([np.convolve(rng.standard_normal(L + 40), rng.standard_normal(6 + 3 * k))[20:20 + L] for k in range(n)])
    order = list(np.roll(np.arange(n), 1))
    est = (s + 0.25 * rng.standard_normal(s.shape) * s.std(axis=1, keepdims=True) + 0.15 * np.roll(s, 1, axis=0))[order]
    s32, e32 = s.astype(np.float32), est.astype(np.float32)
    got = evaluation.bss_eval_sources(torch.tensor(e32, device="cuda"), torch.tensor(s32, device="cuda"), flen=flen)
    want = bsseval.bss_eval_sources(e32.astype(np.float64), s32.astype(np.float64), flen=flen)
    assert list(got[3]) == list(want[3])
    for name, g, w in zip(("SDR", "SIR", "SAR"), got[:3], want[:3]):
        assert np.max(np.abs(g - w)) < 1e-6, (name, g, w)


def test_bss_eval_rejects_host_tensors():
    from deepconvsep_b200 import evaluation
    with pytest.raises(ValueError):
        evaluation.bss_eval_sources(torch.zeros(2, 100), torch.zeros(2, 100))
