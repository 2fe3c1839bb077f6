"""Parity at the REAL configurations of every network family, at the plain north-star tolerance
(per-stem relative L2 <= 1e-4, tests/parity.py -- no whole-stem allowance):

  DSD100          BASELINE configs[1]: one 180 s clip, frameSize 2048, hop 512, tc 30, overlap 25
  Bach10          configs[2]: frameSize 4096 (F = 2049, 214 M parameters), blackmanharris, 10 s
  iKala (pooled)  configs[0]'s network on 10 s incl. a digitally silent segment (max-pool ties everywhere)
  score-informed  configs[4]: frameSize 4096, filters from deepconvsep_b200.score.score_filters on text scores
  stereo / ILD    SURVEY 8(f) row 4: tests/test_gpu_ild.py::test_stereo_medium_clip_strict (15 s)

The oracle runs take a few minutes of host time in total; every case appends its measured errors to
gpurun_out/parity_r2.jsonl."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import dsp, nets, pipeline  # noqa: E402
from parity import strict_check  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(name, arch, F, N, win_name, win_fn, overlap, mix, seed, patcher="standalone", scale=0.3):
    from deepconvsep_b200.engine import Separator
    params = nets.make_synthetic_params(arch, F, seed=seed)
    sep = Separator(params, arch=arch, frame_size=N, hop=512, window=win_name, overlap=overlap, patcher=patcher, feat_size=F)
    extra = None
    if arch == "ikala":      # un-pool routing decisions: see tests/test_gpu_sconv.py::run_case
        got, S, bits = sep.separate_tapped(mix, pool=True)
        want, mag, ph, mm = pipeline.separate(mix, params, arch, frameSize=N, hopSize=512, window=win_fn, overlap=overlap,
                                              patcher=patcher, count_kinks=True, return_spec=True, pool_bits=bits)
        st = pipeline.separate.last_pool_stats
        assert st["disagree_well_conditioned"] == 0 and st["inadmissible"] == 0, st
        assert st["ambiguous"] <= 0.01 * st["windows"], st
        extra = {"pool_windows": st["windows"], "pool_windows_ill_conditioned": st["ambiguous"]}
    else:
        got, S = sep.separate_tapped(mix)
        want, mag, ph, mm = pipeline.separate(mix, params, arch, frameSize=N, hopSize=512, window=win_fn, overlap=overlap,
                                              patcher=patcher, count_kinks=True, return_spec=True)
    kmap = pipeline.separate.last_kink_map
    assert min(np.linalg.norm(w) for w in want) > 0.02 * np.linalg.norm(mix)
    return strict_check(name, got, S, want, mag, ph, mm, kmap, N, 512, win_fn, scale, extra=extra)


def test_dsd100_180s_frame2048():
    mix, _ = pipeline.synth_mixture(180.0, 1000)
    _run("FULL_dsd_N2048_180s", "dsd", 1025, 2048, "hanning", np.hanning, 25, mix, seed=0)


def test_dsd100_60s_frame1024_reference_shape():
    """the frame size every DSD100 script of the reference actually uses"""
    mix, _ = pipeline.synth_mixture(60.0, 1001)
    _run("FULL_dsd_N1024_60s", "dsd", 513, 1024, "hanning", np.hanning, 25, mix, seed=1)


def test_bach10_10s_frame4096():
    mix, _ = pipeline.synth_mixture(10.0, 2000)
    _run("FULL_bach10_N4096_10s", "bach10", 2049, 4096, "blackmanharris", dsp.blackmanharris, 25, mix, seed=2)


def test_ikala_pooled_10s_with_silence():
    mix, _ = pipeline.synth_mixture(10.0, 3000)
    mix[150000:190000] = 0.0          # exact zeros: constant conv1 output -> every max-pool window is a 4-way tie
    _run("FULL_ikala_pool_N1024_10s_silence", "ikala", 513, 1024, "hanning", np.hanning, 20, mix, seed=3)


def test_score_informed_10s_frame4096(tmp_path):
    """filters from the host prelude (score.score_filters) on the text scores of tests/golden/score_golden.npz"""
    from deepconvsep_b200 import score
    from deepconvsep_b200.engine import Separator
    insts = ["bassoon_b", "clarinet_b", "saxophone_b", "violin_b"]
    g = np.load(os.path.join(ROOT, "tests", "golden", "score_golden.npz"))
    for k in insts:
        open(os.path.join(str(tmp_path), k + ".txt"), "wb").write(g["txt_" + k].tobytes())
    N, F = 4096, 2049
    mix, _ = pipeline.synth_mixture(10.0, 4000)
    T = dsp.num_frames(mix.size, 512)
    filters = score.score_filters(str(tmp_path), insts, T, F)
    assert filters.shape == (4, T, F) and filters.dtype == np.float32
    assert (filters.max(axis=(1, 2)) > 0.9).all()           # every instrument has notes inside the 10 s
    params = nets.make_synthetic_params("bach10_score", F, seed=4)
    want, mag, ph, mm = pipeline.separate_score(mix, filters, params, frameSize=N, hopSize=512, window=dsp.blackmanharris,
                                                scale_factor=0.2, overlap=25, count_kinks=True, return_spec=True)
    kmap = pipeline.separate_score.last_kink_map
    sep = Separator(params, arch="bach10_score", frame_size=N, hop=512, window="blackmanharris", overlap=25,
                    patcher="util", scale_factor=0.2, feat_size=F)
    got, S = sep.separate_tapped(mix, filters)
    strict_check("FULL_bach10_score_N4096_10s", got, S, want, mag, ph, mm, kmap, N, 512, dsp.blackmanharris, 0.2)
