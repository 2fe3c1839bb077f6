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
print("done")
