import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dsp, nets, pipeline
from deepconvsep_b200.engine import Separator
arch, F, N, hop, ov, secs, patcher = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6]), sys.argv[7]
win = dsp.blackmanharris if arch.startswith("bach") else np.hanning
params = nets.make_synthetic_params(arch, F, seed=5)
mix, _ = pipeline.synth_mixture(secs, 70 + F)
_, mag, ph, mm = pipeline.separate(mix, params, arch, frameSize=N, hopSize=hop, window=win, overlap=ov, patcher=patcher, return_spec=True)
X = dsp.stft_norm(mix, window=win(N), hopsize=float(hop), nfft=float(N))
T = X.shape[0]
sep = Separator(params, arch=arch, frame_size=N, hop=hop, window=win(N), overlap=ov, patcher=patcher, feat_size=F)
ldf = sep.stft.ldf
magd = torch.zeros((T, ldf), dtype=torch.float32, device="cuda"); Xd = torch.zeros((T, ldf), dtype=torch.complex64, device="cuda")
magd[:, :F] = torch.tensor(mag, device="cuda"); Xd[:, :F] = torch.tensor(X.astype(np.complex64), device="cuda")
S = sep.separate_spec(magd, Xd).cpu().numpy()[:, :, :F].astype(np.complex128)
want = (mm[:, :T] / 0.3) * np.sqrt(N) * np.exp(1j * ph)[None]
if want.shape[1] < T: want = np.concatenate([want, np.zeros((want.shape[0], T - want.shape[1], F))], axis=1)
err = np.abs(S - want)
print("T", T, "P", sep.num_patches(T), "rel", [float(np.linalg.norm(S[s]-want[s])/np.linalg.norm(want[s])) for s in range(S.shape[0])])
ef = (err**2).sum(axis=(0,2)); print("worst frames", np.argsort(-ef)[:12], ef[np.argsort(-ef)[:12]] / (np.abs(want)**2).sum() )
eb = (err**2).sum(axis=(0,1)); print("worst bins", np.argsort(-eb)[:12], eb[np.argsort(-eb)[:12]] / (np.abs(want)**2).sum())
# kink diagnosis at the worst (frame, bin)
from oracle import patch as opatch
t_w = int(np.argmax((err**2).sum(axis=(0,2)))); b_w = int(np.argmax((err**2)[:, t_w].sum(axis=0)))
print("worst (t,b)", t_w, b_w, "err", err[:, t_w, b_w], "want", np.abs(want[:, t_w, b_w]), "|X|", abs(X[t_w, b_w]))
gen = opatch.generate_overlapadd if patcher == "standalone" else opatch.generate_overlapadd_util
batches, n = gen(mag, F, 30, ov, 32)
step = 30 - ov
for k in range(n):
    p = t_w - k * step
    if 0 <= p < 30:
        pre = nets.predict(params, batches[k // 32][k % 32:k % 32 + 1], arch, return_pre=True)[0, :, p, b_w]
        print("patch", k, "p", p, "pre-activations", pre)
