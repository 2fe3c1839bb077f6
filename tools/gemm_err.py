"""Error of the tensor-core GEMM vs K for each accumulator plan (development aid)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch
    from deepconvsep_b200.engine import Context
    ctx = Context(0)
    rng = np.random.default_rng(0)
    out = []
    for K in (32, 128, 512, 750, 2048, 8192):
        M, N = 512, 64
        A = rng.standard_normal((M, K)).astype(np.float32)
        B = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
        ref = A.astype(np.float64) @ B.astype(np.float64)
        e = []
        for eng in (1, 0):
            C = ctx.gemm(torch.tensor(A, device="cuda"), B, None, engine=eng).cpu().numpy().astype(np.float64)
            e.append(np.linalg.norm(C - ref) / np.linalg.norm(ref))
        out.append("K=%d tc=%.2e simt=%.2e" % (K, e[0], e[1]))
    print("acc_mode", os.environ.get("DCS_DEBUG_TC_ACC"), " | ".join(out), flush=True)
else:
    for mode in "012":
        env = dict(os.environ, DCS_DEBUG_TC_ACC=mode)
        subprocess.call([sys.executable, __file__, "run"], env=env)
