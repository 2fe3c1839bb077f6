"""Small end-to-end runs of every architecture (target for compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import nets, pipeline, dsp
from deepconvsep_b200.engine import Separator
mix, _ = pipeline.synth_mixture(1.2, 5)
for arch, F, N, hop, win, ov, patcher in [("dsd", 1025, 2048, 512, "hanning", 25, "standalone"), ("dsd", 513, 1024, 512, "hanning", 25, "util"),
                                           ("dsd", 257, 512, 256, "hanning", 27, "standalone"),   # > 6 patches per frame: FFMA mask kernel
                                           ("bach10", 129, 256, 128, "blackmanharris", 25, "standalone"), ("ikala", 513, 1024, 512, "hanning", 20, "standalone")]:
    params = nets.make_synthetic_params(arch, F, seed=1)
    sep = Separator(params, arch=arch, frame_size=N, hop=hop, window=win, overlap=ov, patcher=patcher, feat_size=F)
    out = sep.separate(mix)
    pcm = sep.separate_pcm16(np.round(mix * 30000).astype(np.int16)) if arch == "dsd" else None
    print(arch, F, ov, out.shape, float(np.abs(out).max()), np.isfinite(out).all(), flush=True)
# round 2: stereo / ILD net (tcgen05 mask kernel with four decoders), score-informed net (window-view conv1, four filter banks
# in the tensor-core K3s), a clip long enough for the persistent A-from-TMEM GEMM to be chosen (>= 8 tiles per SM)
params = nets.make_synthetic_params("dsd_ild", 513, seed=2)
sep = Separator(params, frame_size=1024, hop=512, window="hanning", overlap=25, patcher="util")
st = sep.separate_stereo(np.stack([mix, 0.5 * mix[::-1]], axis=1))
print("dsd_ild", st.shape, np.isfinite(st).all(), flush=True)
F, N, hop = 129, 256, 128
params = nets.make_synthetic_params("bach10_score", F, seed=3)
T = dsp.num_frames(mix.size, hop)
filt = np.full((4, T, F), 0.25, dtype=np.float32)
sep = Separator(params, arch="bach10_score", frame_size=N, hop=hop, window="blackmanharris", overlap=25, patcher="util",
                scale_factor=0.2, feat_size=F)
sc = sep.separate_score(mix, filt)
print("bach10_score", sc.shape, np.isfinite(sc).all(), flush=True)
long_mix, _ = pipeline.synth_mixture(70.0, 6)
params = nets.make_synthetic_params("dsd", 513, seed=4)
sep = Separator(params, frame_size=1024, hop=512, window="hanning", overlap=25)
lo = sep.separate(long_mix)
print("dsd 70 s (persistent GEMM)", lo.shape, np.isfinite(lo).all(), flush=True)
print("done")
