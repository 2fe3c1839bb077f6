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


def _stereo_case(nsrc, L, seed):
    rng = np.random.default_rng(seed)
    s = np.stack([np.stack([np.convolve(rng.standard_normal(L + 40), rng.standard_normal(5 + 2 * k + c))[20:20 + L]
                            for c in range(2)]) for k in range(nsrc)])                       # [nsrc, 2, L]
    est = s + 0.3 * rng.standard_normal(s.shape) * s.std(axis=2, keepdims=True) + 0.2 * np.roll(s, 1, axis=0)
    return s.astype(np.float32), est.astype(np.float32)


def test_bss_eval_windowed_images_matches_oracle_on_device_lags():
    """the variant DSD100 is scored with (evaluation/DSD100_eval_only.m:211-306: `bss_eval(ie, i, win, ove)`, stereo
    images, 512-tap distortion filters, overlapping windows, estimate j against source j, SDR / ISR / SIR / SAR):
    device lags + host algebra against the explicit decomposition of the oracle; the time split of the two
    halves is written to gpurun_out/bsseval_r2.json"""
    import json
    import os
    import time
    from deepconvsep_b200 import evaluation
    from deepconvsep_b200.engine import Context
    nsrc, L, win, ove, flen = 3, 36000, 16000, 8000, 128
    s32, e32 = _stereo_case(nsrc, L, 5)
    ctx = Context(0)
    se, sr = torch.tensor(e32, device="cuda"), torch.tensor(s32, device="cuda")
    got = evaluation.bss_eval_windowed(se, sr, win, ove, flen=flen, ctx=ctx)
    want = bsseval.bss_eval_windowed(e32.astype(np.float64), s32.astype(np.float64), win, ove, flen=flen)
    assert got[0].shape == want[0].shape == (nsrc, len(bsseval.window_starts(L, win, ove))) and got[0].shape[1] >= 3
    for name, g, w in zip(("SDR", "ISR", "SIR", "SAR"), got, want):
        assert np.max(np.abs(g - w)) < 1e-6, (name, g, w)
    # one window at the real filter length (512 taps, the reference's default): where does the time go?
    nsrc2, L2 = 4, 30 * 44100
    s2, e2 = _stereo_case(nsrc2, L2, 6)
    se2, sr2 = torch.tensor(e2, device="cuda"), torch.tensor(s2, device="cuda")
    kinds, idx = evaluation.image_pair_list(nsrc2, 2)
    sig = {"ss": (sr2, sr2), "se": (sr2, se2), "ee": (se2, se2)}
    pairs = [(sig[k][0][a // 2, a % 2], sig[k][1][b // 2, b % 2]) for k, a, b in kinds]
    evaluation.xcorr_lags(ctx, pairs[:4], L2, 512)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    R = evaluation.xcorr_lags(ctx, pairs, L2, 512)
    t_lags = time.perf_counter() - t0
    t0 = time.perf_counter()
    r_host = evaluation.images_from_lags(R, idx, nsrc2, 2, 512)                      # numpy on the host
    t_host = time.perf_counter() - t0
    evaluation.images_from_lags(R, idx, nsrc2, 2, 512, device=se2.device)            # warm-up (cuSOLVER handles)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = evaluation.images_from_lags(R, idx, nsrc2, 2, 512, device=se2.device)        # float64 Cholesky on the GPU
    t_dev = time.perf_counter() - t0
    assert all(np.isfinite(x).all() for x in r)
    for a, b in zip(r, r_host):
        assert np.max(np.abs(a - b)) < 1e-6
    # algorithmic work of the lag kernel: every pair reads its two signals once per 4-lag group
    rec = {"case": "bss_eval_images, 4 stereo sources, one 30 s window, 512 taps", "pairs": len(pairs),
           "device_lags_s": t_lags, "host_gram_and_solve_s": t_host, "gram_on_host_plus_device_cholesky_s": t_dev,
           "lag_macs": len(pairs) * 1023 * float(L2), "lag_fp64_gflops": 2 * len(pairs) * 1023 * float(L2) / t_lags / 1e9}
    os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "bsseval_r2.json"), "w") as f:
        json.dump(rec, f)
