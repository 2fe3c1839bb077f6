"""GPU parity of the stereo / ILD DSD100 variant (examples/dsd100_2ch_ILD/trainCNN_ILD_DSD100.py,
SURVEY.md 8(f) row 4) against the float64 oracle: per (source, channel) relative L2 <= 1e-4."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import nets, pipeline  # noqa: E402
from parity import strict_check, TOL  # noqa: E402


def run_strict(name, sep, params, audio, N, hop):
    """per channel j: stems [4, L] against the oracle with channel j's flagged bins taken out (tests/parity.py)"""
    want, mag, phs, mms = pipeline.separate_stereo(audio, params, frameSize=N, hopSize=hop, count_kinks=True, return_spec=True)
    kmap = pipeline.separate_stereo.last_kink_map
    got, S = sep.separate_tapped(audio)          # got [L, 4, 2]; S planes ordered (source, channel)
    assert got.shape == want.shape == (audio.shape[0], 4, 2) and got.dtype == np.float32
    for j in range(2):
        strict_check("%s_ch%d" % (name, j), np.ascontiguousarray(got[:, :, j].T), S[j::2], np.ascontiguousarray(want[:, :, j].T),
                     mag[j], phs[j], mms[j], kmap[j], N, hop, np.hanning, 0.3)
    return got, want


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def stereo_clip(seconds, seed):
    mix, _ = pipeline.synth_mixture(seconds, seed)
    other, _ = pipeline.synth_mixture(seconds, seed + 1)
    return np.stack([0.7 * mix + 0.3 * other, 0.4 * mix + 0.6 * np.roll(other, 11)], axis=1)


@pytest.mark.parametrize("N,seconds,seed", [(1024, 3.0, 1029), (2048, 2.5, 2050), (512, 1.2, 516)])
def test_stereo_matches_oracle(N, seconds, seed):
    from deepconvsep_b200.engine import Separator
    F, hop = N // 2 + 1, min(512, N // 2)
    params = nets.make_synthetic_params("dsd_ild", F, seed=seed)
    sep = Separator(params, frame_size=N, hop=hop, window="hanning", overlap=25, patcher="util")
    assert sep.model.arch == "dsd_ild" and sep.nsrc == 4
    audio = stereo_clip(seconds, 300 + N)
    got, want = run_strict("dsd_ild_N%d_%gs" % (N, seconds), sep, params, audio, N, hop)
    for i in range(4):
        for j in range(2):
            # every (source, channel) must carry energy, or the comparison exercises nothing
            assert np.linalg.norm(want[:, i, j]) > 0.01 * np.linalg.norm(audio[:, j])
    # device planes: (source, channel) order, same bits as the host-buffer call
    d = sep.separate_stereo(torch.tensor(np.ascontiguousarray(audio.T), dtype=torch.float32, device="cuda"))
    assert np.array_equal(d.cpu().numpy().reshape(4, 2, -1).transpose(2, 0, 1), got)


def test_stereo_medium_clip_strict():
    """15 s clip"""
    from deepconvsep_b200.engine import Separator
    N = 1024
    params = nets.make_synthetic_params("dsd_ild", N // 2 + 1, seed=77)
    sep = Separator(params, frame_size=N, hop=512, window="hanning", overlap=25, patcher="util")
    audio = stereo_clip(15.0, 4321)
    run_strict("dsd_ild_N1024_15s", sep, params, audio, N, 512)


def test_stereo_silent_channel():
    from deepconvsep_b200.engine import Separator
    N = 1024
    params = nets.make_synthetic_params("dsd_ild", N // 2 + 1, seed=5)
    sep = Separator(params, frame_size=N, hop=512, window="hanning", overlap=25, patcher="util")
    audio = stereo_clip(2.0, 9)
    audio[:, 1] = 0.0
    got = sep.separate_stereo(audio)
    assert np.all(got[:, :, 1] == 0) and np.isfinite(got).all()
