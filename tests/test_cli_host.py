"""The stand-alone scripts' host side without a GPU: a stand-in for the CUDA Separator (scales its input per source)
behind `_common.get_separator`; checked are the things the reference scripts fix on the host -- which arguments
`train_auto` receives from `main`, the decode / downmix conventions, the int16 truncation, the output file names
(examples/dsd100/separate_dsd.py:243,309,332; ikala/separate_ikala.py:229,253-254,275; bach10/separate_bach10.py:236,302,325)."""
import numpy as np
import pytest
import scipy.io.wavfile
from types import SimpleNamespace

from deepconvsep_b200.examples import _common
from deepconvsep_b200.models import FAMILY_DEFAULTS

GAINS = (0.5, 0.25, 0.125, 0.0625)


class FakeSeparator(object):
    def __init__(self, family, frame_size, hop, overlap):
        self.model = SimpleNamespace(arch=family, tc=30)
        self.frame_size, self.hop, self.overlap = frame_size, hop, overlap
        self.sources = FAMILY_DEFAULTS[family]["sources"]
        self.nsrc = len(self.sources)

    def separate(self, audio):
        return np.stack([np.asarray(audio, dtype=np.float32) * np.float32(g) for g in GAINS[:self.nsrc]])

    def separate_pcm16(self, pcm, downmix=1):
        p = np.asarray(pcm)
        mono = p.astype(np.float32) / np.float32(32767) if p.ndim == 1 else \
            (p[:, 0].astype(np.float32) + p[:, 1].astype(np.float32)) / np.float32(2 * 32767)
        return np.stack([(mono * np.float32(g) * np.float32(32767)).astype(np.int16) for g in GAINS[:self.nsrc]])


@pytest.fixture
def seen(monkeypatch):
    calls = []

    def fake(model, arch, frame_size, hop, window, scale_factor, time_context, overlap, feat_size, device=0, slot=0):
        calls.append(dict(model=model, arch=arch, frame_size=frame_size, hop=hop, window=window, scale_factor=scale_factor,
                          time_context=time_context, overlap=overlap, feat_size=feat_size, device=device, slot=slot))
        family = arch or "ikala"
        return FakeSeparator(family, frame_size, hop, overlap)
    monkeypatch.setattr(_common, "get_separator", fake)
    return calls


def _wav(tmp_path, name, seconds=1.0, channels=2, seed=0):
    rng = np.random.default_rng(seed)
    shape = (int(44100 * seconds), channels) if channels > 1 else (int(44100 * seconds),)
    pcm = (rng.uniform(-0.4, 0.4, size=shape) * 32767).astype(np.int16)
    p = tmp_path / name
    scipy.io.wavfile.write(str(p), 44100, pcm)
    return p, pcm


def test_dsd_script(tmp_path, seen):
    from deepconvsep_b200.examples.dsd100 import separate_dsd
    wav, pcm = _wav(tmp_path, "mix.wav")
    out = tmp_path / "o"
    out.mkdir()
    separate_dsd.main(["-i", str(wav), "-o", str(out), "-m", "m.pkl"])
    c = seen[-1]
    assert (c["arch"], c["frame_size"], c["hop"], c["overlap"], c["time_context"], c["scale_factor"], c["feat_size"]) == \
        ("dsd", 1024, 512, 25, 30, 0.3, 513)                                      # separate_dsd.py:332
    assert sorted(f.name for f in out.iterdir()) == ["bass.wav", "drums.wav", "other.wav", "vocals.wav"]
    sr, v = scipy.io.wavfile.read(str(out / "vocals.wav"))
    assert sr == 44100 and v.dtype == np.int16 and v.shape == (len(pcm),)
    # hiphopss is the same script under another name (examples/hiphopss/separate_hhds.py is a byte copy in the reference)
    from deepconvsep_b200.examples.hiphopss import separate_hhds
    assert separate_hhds.train_auto is separate_dsd.train_auto and separate_hhds.main is separate_dsd.main


def test_ikala_script_sums_the_channels_and_names_its_outputs(tmp_path, seen):
    from deepconvsep_b200.examples.ikala import separate_ikala
    wav, pcm = _wav(tmp_path, "song7.wav", seed=1)
    out = tmp_path / "o"
    out.mkdir()
    separate_ikala.main(["-i", str(wav), "-o", str(out), "-m", "m.pkl"])
    c = seen[-1]
    assert (c["arch"], c["frame_size"], c["overlap"], c["feat_size"]) == (None, 1024, 20, 513)    # separate_ikala.py:275
    assert sorted(f.name for f in out.iterdir()) == ["song7-music.wav", "song7-voice.wav"]          # :253-254
    mono = pcm[:, 0] / 32767.0 + pcm[:, 1] / 32767.0                                              # L + R, not halved (:229)
    sr, v = scipy.io.wavfile.read(str(out / "song7-voice.wav"))
    want = ((mono.astype(np.float32) * np.float32(0.5)).astype(np.float64) * 32767).astype(np.int16)
    assert np.array_equal(v, want)


def test_bach10_script(tmp_path, seen):
    from deepconvsep_b200.examples.bach10 import separate_bach10
    wav, pcm = _wav(tmp_path, "01-AchGottundHerr.wav", channels=1, seed=2)
    out = tmp_path / "o"
    out.mkdir()
    separate_bach10.main(["-i", str(wav), "-o", str(out), "-m", "m.pkl"])
    c = seen[-1]
    assert (c["arch"], c["frame_size"], c["hop"], c["overlap"], c["feat_size"]) == ("bach10", 4096, 512, 25, 2049)   # :325
    assert c["window"] == "blackmanharris"
    names = sorted(f.name for f in out.iterdir())
    assert names == ["01-AchGottundHerr_%s.wav" % s for s in ("bassoon", "clarinet", "saxphone", "violin")]   # the reference's spelling


def test_wrong_sample_rate_is_reported_not_separated(tmp_path, seen, capsys):
    from deepconvsep_b200.examples.dsd100 import separate_dsd
    p = tmp_path / "lo.wav"
    scipy.io.wavfile.write(str(p), 22050, np.zeros(2205, dtype=np.int16))
    out = tmp_path / "o"
    out.mkdir()
    assert separate_dsd.train_auto(str(p), str(out), "m.pkl") is None
    assert "Sample rate is not 44100" in capsys.readouterr().out                  # separate_dsd.py:313
    assert list(out.iterdir()) == [] and seen == []


def test_float_wav_keeps_the_reference_normalisation(tmp_path, seen):
    """float wavs are divided by finfo.max (separate_dsd.py:277-280): practically silent output, as in the reference"""
    from deepconvsep_b200.examples.dsd100 import separate_dsd
    p = tmp_path / "f.wav"
    scipy.io.wavfile.write(str(p), 44100, (np.random.default_rng(3).uniform(-0.5, 0.5, 4410)).astype(np.float32))
    out = tmp_path / "o"
    out.mkdir()
    separate_dsd.train_auto(str(p), str(out), "m.pkl")
    sr, v = scipy.io.wavfile.read(str(out / "vocals.wav"))
    assert v.dtype == np.int16 and not v.any()
