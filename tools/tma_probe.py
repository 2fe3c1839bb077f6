"""TMA-fed GEMM bring-up (development aid): accuracy over shapes, then the DSD100 pipeline per mode
(DCS_DEBUG_TMA = 0 register-staged, 1 TMA, 2 TMA + rewritten high plane) with stage times and the
output difference against mode 0."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch
    from deepconvsep_b200.engine import Context, Separator

    def synth_params(F, seed=0):
        rng = np.random.default_rng(seed)
        shp = [(50,1,1,F),(50,),(50,),(50,50,15,1),(50,),(50,),(800,128),(128,),(128,800),(800,),(128,800),(800,),(128,800),(800,),(4,)]
        out = []
        for s in shp:
            a = np.sqrt(6.0/((s[0]+s[1])*s[2]*s[3])) if len(s)==4 else (np.sqrt(6.0/(s[0]+s[1])) if len(s)==2 else 0.1)
            out.append(rng.uniform(-a,a,size=s).astype(np.float32))
        return out
    tag = "TMA=%s mask=%s stages=%s" % (os.environ.get("DCS_DEBUG_TMA", "1"), os.environ.get("DCS_DEBUG_TMA_MASK", "15"), os.environ.get("DCS_DEBUG_TMA_STAGES", "2") + " istft_waves=" + os.environ.get("DCS_DEBUG_ISTFT_WAVES", "1"))
    if sys.argv[1] == "gemm":
        ctx = Context(0)
        rng = np.random.default_rng(0)
        out = []
        for (M, N, K, lda) in ((512, 64, 750, 752), (1000, 50, 1025, 1028), (129, 30, 100, 100), (3100, 2400, 128, 128),
                               (77, 200, 40, 64), (4096, 50, 32, 32), (300, 64, 2048, 2048)):
            Afull = torch.tensor(rng.standard_normal((M, lda)).astype(np.float32), device="cuda")
            A = Afull[:, :K]
            B = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
            bias = rng.standard_normal(N).astype(np.float32)
            ref = A.cpu().numpy().astype(np.float64) @ B.astype(np.float64) + bias
            C = ctx.gemm(A, B, bias, engine=1).cpu().numpy().astype(np.float64)
            out.append("%dx%dx%d %.1e" % (M, N, K, np.linalg.norm(C - ref) / np.linalg.norm(ref)))
        print(tag, "gemm rel err:", " | ".join(out), flush=True)
    else:
        sep = Separator(synth_params(1025), frame_size=2048, hop=512, window="hanning", overlap=25)
        L = int(sys.argv[2]) if len(sys.argv) > 2 else 7938000
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        x = (torch.rand(L, device="cuda", generator=g) - 0.5) * 0.4
        out = torch.empty((4, L), device="cuda")
        for _ in range(3): sep.separate_device(x, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): sep.separate_device(x, out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        sep.ctx.profile(True); sep.separate_device(x, out); torch.cuda.synchronize()
        st = " ".join("%s=%.3f" % (k.replace("_gemm", ""), v) for k, v in sep.ctx.profile_read())
        ref_path = "/tmp/tma_probe_ref.pt"
        if os.environ.get("DCS_DEBUG_TMA") == "0":
            torch.save(out.cpu(), ref_path); diff = "ref"
        elif os.path.exists(ref_path) and L == 7938000:
            r = torch.load(ref_path).cuda()
            diff = "rel diff vs mode0 %.2e" % float((out - r).norm() / r.norm())
        else:
            diff = "no ref"
        print(tag, "N=2048 %.3f ms | %s | %s | finite %s" % (ms, st, diff, bool(torch.isfinite(out).all())), flush=True)
else:
    def run(env, what, *args):
        try:
            return subprocess.call([sys.executable, __file__, what] + list(args), env=dict(os.environ, **env), timeout=150)
        except subprocess.TimeoutExpired:
            print("TIMEOUT", env, what, flush=True)
            return -1
    run({"DCS_DEBUG_TMA": "0"}, "pipe")
    run({}, "pipe")
    run({"DCS_DEBUG_ISTFT_WAVES": "2"}, "pipe")
