"""Host score prelude (deepconvsep_b200/score.py) against vectors produced by the reference's own
util.getMidiNum / expandMidi / str2midi and LargeDatasetMask2.filterSpec (tests/golden/make_golden.py)."""
import os
import numpy as np
import pytest

from deepconvsep_b200 import score

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTS = ["bassoon_b", "clarinet_b", "saxophone_b", "violin_b"]


@pytest.fixture(scope="module")
def sg(tmp_path_factory):
    g = np.load(os.path.join(ROOT, "tests", "golden", "score_golden.npz"))
    d = tmp_path_factory.mktemp("scores")
    for k in INSTS:
        open(os.path.join(str(d), k + ".txt"), "wb").write(g["txt_" + k].tobytes())
    return g, str(d)


def test_str2midi_and_slices(sg):
    g, _ = sg
    got = [score.str2midi(s) for s in ["C3", "Bb4", "F#3", "C#5", "A4", "Ebb2", "Gx6"]]
    np.testing.assert_array_equal(np.array(got, dtype=np.float64), g["str2midi"])
    assert score.str2midi(b"A4") == 69 and np.isnan(score.str2midi("?"))
    assert score.slicefft_slices(0, 4096) == []
    sl = score.slicefft_slices(57, 4096, interval=50, nharmonics=20)
    assert all(s.stop <= 2049 for s in sl) and all(a.stop < b.start for a, b in zip(sl, sl[1:]))


def test_getMidiNum_expandMidi_filterSpec_match_reference(sg):
    g, d = sg
    nframes = int(g["nframes"])
    melody = np.zeros_like(g["melody"])
    for i, inst in enumerate(INSTS):
        assert score.getMidiNum(inst, d, 0, 40.0) == int(g["num_%d" % i])
        tmp = score.expandMidi(inst, d, 0, 40.0, 50, 440, 20, 44100, 512, 4096, 0.2, 0.2, nframes, 0.5)
        np.testing.assert_array_equal(tmp, g["exp_%d" % i])
        melody[i, :tmp.shape[0]] = tmp
    np.testing.assert_array_equal(melody, g["melody"])
    mask = score.filterSpec(np.zeros((nframes, 2049), dtype=np.float32), melody, 0, nframes)
    assert mask.dtype == np.float32
    np.testing.assert_array_equal(mask, g["mask"])
    # the normalised filters sum to one in every bin, so the sum of the four input channels is the mixture
    np.testing.assert_allclose(mask.reshape(nframes, 4, 2049).sum(axis=1), 1.0, rtol=0, atol=3e-7)
    planes = score.score_filters(d, INSTS, nframes, 2049)
    np.testing.assert_array_equal(planes, mask.reshape(nframes, 4, 2049).transpose(1, 0, 2))
