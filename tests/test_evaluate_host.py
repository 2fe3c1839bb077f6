"""Host glue of deepconvsep_b200.evaluate (the reference's MATLAB evaluation drivers): file conventions,
gain normalisation, NSDR logic, window bookkeeping and the .mat layout -- with the metric itself
replaced by the oracle restatement (no GPU here; the device lags are covered by tests/test_gpu_bsseval.py
and the lag algebra by tests/test_evaluation_host.py)."""
import os
import numpy as np
import pytest
import scipy.io
import scipy.io.wavfile

from oracle import bsseval
from deepconvsep_b200 import evaluate, evaluation

FS = 44100


@pytest.fixture
def cpu_metric(monkeypatch):
    monkeypatch.setattr(evaluate, "DEVICE", "cpu")
    monkeypatch.setattr(evaluation, "bss_eval_sources",
                        lambda est, ref, flen=512, **kw: bsseval.bss_eval_sources(est.numpy().astype(np.float64),
                                                                                  ref.numpy().astype(np.float64), flen))
    monkeypatch.setattr(evaluation, "bss_eval_windowed",
                        lambda est, ref, win, ove, flen=512, **kw: bsseval.bss_eval_windowed(
                            est.numpy().astype(np.float64), ref.numpy().astype(np.float64), win, ove, flen))


def _wav(path, x):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    scipy.io.wavfile.write(path, FS, (np.clip(x, -1, 1) * 32767).astype(np.int16))


def _noise(rng, n, k):
    return 0.1 * np.convolve(rng.standard_normal(n + 16), rng.standard_normal(3 + k))[8:8 + n]


def test_ikala_driver(tmp_path, cpu_metric):
    rng = np.random.default_rng(0)
    n = 6000
    voice, music = _noise(rng, n, 0), _noise(rng, n, 2)
    root = str(tmp_path)
    _wav(os.path.join(root, "Wavfile", "a.wav"), np.stack([music, voice], axis=1))     # ch 1 music, ch 2 voice
    ev = voice + 0.2 * music + 0.02 * rng.standard_normal(n)
    ek = music + 0.1 * voice + 0.02 * rng.standard_normal(n)
    _wav(os.path.join(root, "output", "m1", "a-voice.wav"), ev)
    _wav(os.path.join(root, "output", "m1", "a-music.wav"), np.concatenate([ek, np.zeros(50)]))   # longer: truncated
    done = evaluate.evaluate_ikala(root, "m1", flen=16)
    assert done == [os.path.join(root, "measures", "test_m1", "a.mat")]
    m = scipy.io.loadmat(done[0])
    assert set(k for k in m if not k.startswith("__")) == {"SDR", "SIR", "SAR", "NSDR", "NSIR", "NSAR"}
    assert m["SDR"].shape == (2, 1)
    # the estimates are better than the mixture: positive normalised SDR; and the definition holds
    assert m["NSDR"].min() > 3
    q = lambda x: (np.clip(x, -1, 1) * 32767).astype(np.int16) / 32767.0
    v, k = q(voice), q(music)
    mix = (v + k) / 2
    base = bsseval.bss_eval_sources(np.stack([mix, mix]) / np.linalg.norm(2 * mix), np.stack([v, k]) / np.linalg.norm(v + k), 16)
    assert np.allclose(m["SDR"][:, 0] - m["NSDR"][:, 0], base[0], atol=1e-9)
    assert evaluate.evaluate_ikala(root, "m1", flen=16) == []          # existing results are kept


def test_dsd100_driver(tmp_path, cpu_metric):
    rng = np.random.default_rng(1)
    n = FS // 2                                                        # 0.5 s, windows of 0.2 s every 0.1 s
    ds, es = str(tmp_path / "DSD100"), str(tmp_path / "est")
    song = "001 - x"
    refs = {s: np.stack([_noise(rng, n, i), _noise(rng, n, i + 1)], axis=1) for i, s in enumerate(evaluate.DSD_SOURCES)}
    for s, x in refs.items():
        _wav(os.path.join(ds, "Sources", "Test", song, s + ".wav"), x)
    names = dict(zip(evaluate.DSD_SOURCES, evaluate.DSD_ESTIMATE_FILES))
    for s, x in refs.items():
        if s == "bass":                                                # a mono, shorter estimate: duplicated to stereo,
            e = (x[:, 0] + 0.005 * rng.standard_normal(n))[: n - 100]  # and every signal is cut to its length
        else:
            e = x + 0.005 * rng.standard_normal(x.shape)
        _wav(os.path.join(es, "Test", song, names[s] + ".wav"), e)
    done = evaluate.evaluate_dsd100(ds, es, win_s=0.2, hop_s=0.1, flen=8)
    assert done == [os.path.join(es, "Test", song + "_results.mat")]
    r = scipy.io.loadmat(done[0], squeeze_me=True, struct_as_record=False)["results"]
    assert r.name == song
    nwin = len(evaluation.window_starts(n - 100, int(0.2 * FS), int(0.1 * FS)))     # shortest estimate rules
    for s in evaluate.DSD_SOURCES + ["accompaniment"]:
        for mname in ("sdr", "isr", "sir", "sar"):
            v = np.atleast_1d(getattr(getattr(r, s), mname))
            assert v.shape == (nwin,) and np.isfinite(v).all()
    assert np.median(r.vocals.sdr) > 10 and np.median(r.accompaniment.sdr) < np.median(r.vocals.sdr)
