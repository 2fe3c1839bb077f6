"""Run the DSD100 pipeline on one 180 s clip a few times (target for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_params, synth_clip_device
from deepconvsep_b200.engine import Separator
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sep = Separator(synthetic_params(N // 2 + 1, 0), frame_size=N, hop=512, window="hanning", overlap=25)
x = synth_clip_device(7938000, 1000, torch.device("cuda", 0))
out = torch.empty((4, x.numel()), device="cuda")
for _ in range(reps):
    sep.separate_device(x, out)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
