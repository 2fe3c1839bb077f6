"""Quick device-side timing of the DSD100 pipeline (development aid, not the bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepconvsep_b200.engine import Separator

def synth_params(F, seed=0):
    rng = np.random.default_rng(seed)
    shp = [(50,1,1,F),(50,),(50,),(50,50,15,1),(50,),(50,),(800,128),(128,),(128,800),(800,),(128,800),(800,),(128,800),(800,),(4,)]
    out=[]
    for s in shp:
        a = np.sqrt(6.0/((s[0]+s[1])*s[2]*s[3])) if len(s)==4 else (np.sqrt(6.0/(s[0]+s[1])) if len(s)==2 else 0.1)
        out.append(rng.uniform(-a,a,size=s).astype(np.float32))
    return out

for N in (1024, 2048):
    sep = Separator(synth_params(N//2+1), frame_size=N, hop=512, window="hanning", overlap=25)
    x = (torch.rand(7938000, device="cuda") - 0.5) * 0.4
    out = torch.empty((4, x.numel()), device="cuda")
    for _ in range(3): sep.separate_device(x, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): sep.separate_device(x, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    print("N=%d: %.3f ms per 180 s clip -> %.0fx real time; workspace %.2f GB" % (N, ms, 180.0/(ms*1e-3), sep.ctx.workspace_bytes()/1e9), flush=True)
