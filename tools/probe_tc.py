import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from bench import synthetic_params, synth_clip_device
    from deepconvsep_b200.engine import Separator
    sep = Separator(synthetic_params(1025, 0), frame_size=2048, hop=512, window="hanning", overlap=25)
    x = synth_clip_device(7938000, 1000, torch.device("cuda", 0)); out = torch.empty((4, x.numel()), device="cuda")
    for _ in range(2): sep.separate_device(x, out)
    torch.cuda.synchronize(); sep.ctx.profile(True)
    for _ in range(3): sep.separate_device(x, out)
    torch.cuda.synchronize()
    acc = {}
    for k, v in sep.ctx.profile_read(): acc.setdefault(k, []).append(v)
    print("skip=%s " % os.environ.get("DCS_DEBUG_TC_SKIP", "0") + " ".join("%s=%.3f" % (k.replace("_gemm", ""), sum(v)/len(v)) for k, v in acc.items() if "gemm" in k), flush=True)
else:
    for m in "01234567":
        subprocess.call([sys.executable, __file__, "run"], env=dict(os.environ, DCS_DEBUG_TC_SKIP=m))
