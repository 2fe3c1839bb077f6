"""The dataset runner's host side without a GPU (deepconvsep_b200/runner.py; trainers' `if not skip_sep:` branch,
examples/dsd100/trainCNN.py:285-335, ikala/trainCNN.py:246-285, dsd100_2ch_ILD/trainCNN_ILD_DSD100.py:291-343):
directory conventions, downmix per family, the input's bit depth on the way out, trainer settings handed to the
separator, songs of a shard longest first, two ranks covering the dataset once."""
import os
import numpy as np
import scipy.io.wavfile
from types import SimpleNamespace

from deepconvsep_b200 import runner
from deepconvsep_b200.models import FAMILY_DEFAULTS


def _song(path, seconds, channels=2, dtype=np.int16, seed=0):
    rng = np.random.default_rng(seed)
    n = int(44100 * seconds)
    x = rng.uniform(-0.4, 0.4, size=(n, channels) if channels > 1 else (n,))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    scipy.io.wavfile.write(path, 44100, (x * np.iinfo(dtype).max).astype(dtype))


def _fake_separator(log):
    class Fake(object):
        def __init__(self, params, arch=None, frame_size=None, hop=None, window=None, scale_factor=0.3, time_context=None,
                     overlap=None, patcher="standalone", device=0, feat_size=None):
            family = arch or "ikala"
            log.append(("init", dict(arch=arch, frame_size=frame_size, hop=hop, window=window, overlap=overlap, patcher=patcher,
                                     device=device, feat_size=feat_size, scale_factor=scale_factor)))
            self.sources = FAMILY_DEFAULTS[family]["sources"]
            self.nsrc = len(self.sources)
            self.model = SimpleNamespace(arch=family, tc=30)

        def separate(self, audio):
            log.append(("separate", len(audio), float(np.abs(audio).max())))
            return np.stack([np.asarray(audio, dtype=np.float32) / (s + 1) for s in range(self.nsrc)])

        def separate_stereo(self, audio):
            log.append(("stereo", audio.shape))
            a = np.asarray(audio, dtype=np.float32)
            return np.stack([a / (s + 1) for s in range(self.nsrc)], axis=1)           # [L, nsrc, 2]
    return Fake


def test_dsd_layout_downmix_bit_depth_and_order(tmp_path, monkeypatch):
    log = []
    monkeypatch.setattr(runner, "Separator", _fake_separator(log))
    db, out = tmp_path / "Mixtures", tmp_path / "out"
    _song(str(db / "Dev" / "051 - A" / "mixture.wav"), 0.5, seed=1)
    _song(str(db / "Test" / "005 - B" / "mixture.wav"), 1.5, seed=2)
    _song(str(db / "Test" / "007 - C" / "mixture.wav"), 1.0, dtype=np.int32, seed=3)
    (db / "Test" / ".DS_Store").mkdir()
    secs, njobs = runner.separate_dataset("dsd", str(db), str(out), model=[np.zeros(1)])
    assert njobs == 3 and abs(secs - 3.0) < 1e-3
    init = log[0][1]
    assert (init["arch"], init["frame_size"], init["hop"], init["window"], init["overlap"], init["patcher"], init["feat_size"]) == \
        ("dsd", 1024, 512, "blackmanharris", 25, "util", 513)                       # dsd100/trainCNN.py:399,431
    assert [e[1] for e in log if e[0] == "separate"] == [66150, 44100, 22050]      # longest first
    for sub, song, dt in (("Dev", "051 - A", np.int16), ("Test", "005 - B", np.int16), ("Test", "007 - C", np.int32)):
        for i, s in enumerate(("vocals", "bass", "drums", "other")):
            sr, y = scipy.io.wavfile.read(str(out / sub / song / (s + ".wav")))
            assert sr == 44100 and y.dtype == dt and y.ndim == 1                    # the input's bit depth (util.py:56-58)
    sr, mix = scipy.io.wavfile.read(str(db / "Test" / "005 - B" / "mixture.wav"))
    sr, voc = scipy.io.wavfile.read(str(out / "Test" / "005 - B" / "vocals.wav"))
    mono = (mix[:, 0] / 32767.0 + mix[:, 1] / 32767.0) / 2                          # dsd100/trainCNN.py:304
    assert np.array_equal(voc, (mono.astype(np.float32).astype(np.float64) * 32767).astype(np.int16))


def test_ikala_sums_the_channels_and_two_ranks_cover_the_dataset_once(tmp_path, monkeypatch):
    log = []
    monkeypatch.setattr(runner, "Separator", _fake_separator(log))
    db = tmp_path / "Wavfile"
    for k, secs in enumerate((0.4, 0.9, 0.6, 0.2, 0.7)):
        _song(str(db / ("1000%d_verse.wav" % k)), secs, seed=10 + k)
    done = []
    for rank in range(2):
        out = tmp_path / "out"
        secs, njobs = runner.separate_dataset("ikala", str(db), str(out), model=[np.zeros(1)], rank=rank, world_size=2, device=rank)
        assert njobs == 5
        done.append(secs)
    assert abs(sum(done) - 2.8) < 1e-3 and min(done) > 0.9                          # balanced: 1.5 s vs 1.3 s
    inits = [e[1] for e in log if e[0] == "init"]
    assert [i["device"] for i in inits] == [0, 1] and inits[0]["arch"] is None and inits[0]["overlap"] == 20   # ikala/trainCNN.py:382
    names = sorted(os.listdir(str(tmp_path / "out")))
    assert names == sorted("1000%d_verse-%s.wav" % (k, s) for k in range(5) for s in ("voice", "music"))
    sr, mix = scipy.io.wavfile.read(str(db / "10001_verse.wav"))
    sr, voice = scipy.io.wavfile.read(str(tmp_path / "out" / "10001_verse-voice.wav"))
    summed = mix[:, 0] / 32767.0 + mix[:, 1] / 32767.0                              # L + R, not halved (ikala/trainCNN.py:255)
    assert np.array_equal(voice, (summed.astype(np.float32).astype(np.float64) * 32767).astype(np.int16))


def test_stereo_ild_layout(tmp_path, monkeypatch):
    log = []
    monkeypatch.setattr(runner, "Separator", _fake_separator(log))
    root, out = tmp_path / "DSD100", tmp_path / "out"
    _song(str(root / "Mixtures" / "Test" / "002 - X" / "mixture.wav"), 0.3, seed=5)
    secs, njobs = runner.separate_dataset("dsd_ild", str(root), str(out), model=[np.zeros(1)])
    assert njobs == 1 and [e[0] for e in log] == ["init", "stereo"]
    init = log[0][1]
    assert (init["window"], init["overlap"], init["frame_size"]) == ("hanning", 25, 1024)     # trainCNN_ILD_DSD100.py:438-441,487
    for s in ("vocals", "bass", "drums", "other"):
        sr, y = scipy.io.wavfile.read(str(out / "Sources" / "Test" / "002 - X" / (s + ".wav")))
        assert y.shape == (13230, 2) and y.dtype == np.int16                                   # stereo stems (:329-343)
