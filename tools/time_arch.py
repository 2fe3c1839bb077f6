"""Device-side timing + stage table for a non-DSD architecture (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepconvsep_b200.engine import Separator
from bench import synth_clip_device
arch = sys.argv[1]; secs = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
cfg = {"bach10": (2049, 4096, "blackmanharris", 25, 4), "ikala": (513, 1024, "hanning", 20, 2)}[arch]
F, N, win, ov, nsrc = cfg
def shapes():
    if arch == "bach10":
        J = (F - 30) // 4 + 1; flat = 30 * 11 * J
        return [(30,1,1,30),(30,),(30,),(30,30,20,1),(30,),(30,),(flat,256),(256,)] + [(256,flat),(flat,)]*4 + [(4,)]
    J = (F - 30) // 3 + 1; WP = J // 4; flat = 30 * 21 * (WP - 19)
    return [(30,1,1,30),(30,),(30,),(30,30,10,20),(30,),(30,),(flat,256),(256,)] + [(256,flat),(flat,)]*2 + [(2,)]
rng = np.random.default_rng(0); params = []
sh = shapes()
for i, s in enumerate(sh):
    a = np.sqrt(6.0/((s[0]+s[1])*s[2]*s[3])) if len(s)==4 else (np.sqrt(6.0/(s[0]+s[1])) if len(s)==2 else (0.002 if i==len(sh)-1 else 0.1))
    params.append((rng.random(s, dtype=np.float32)*2-1)*np.float32(a))
t0 = time.time()
sep = Separator(params, arch=arch, frame_size=N, hop=512, window=win, overlap=ov, feat_size=F)
print("model upload %.1f s" % (time.time()-t0), flush=True)
L = int(secs*44100)
x = synth_clip_device(L, 1000, torch.device("cuda", 0)); out = torch.empty((nsrc, L), device="cuda")
for _ in range(2): sep.separate_device(x, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): sep.separate_device(x, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/3
print("%s: %.2f ms per %.0f s clip -> %.0fx real time; workspace %.2f GB; finite %s" % (arch, ms, secs, secs/(ms*1e-3), sep.ctx.workspace_bytes()/1e9, bool(torch.isfinite(out).all())), flush=True)
sep.ctx.profile(True); sep.separate_device(x, out); torch.cuda.synchronize()
for k, v in sep.ctx.profile_read(): print("  %-26s %8.3f ms" % (k, v))
