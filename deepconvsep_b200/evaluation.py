"""BSS-Eval 3.0 `bss_eval_sources` (evaluation/bss_eval/bss_eval_sources.m; used by
evaluation/evaluate_SS_iKala.m:58-59 and evaluation/Bach10_eval_only.m:94) with the O(L) work on
the GPU -- SURVEY.md 8(f) row 3.

The reference decomposes each estimate into s_true + e_spat + e_interf + e_artif with 512-tap
least-squares projections whose normal equations are made of correlation lags |m| < 512
(:110-159).  Here

  * libdcs computes every needed lag in float64 on the device (`dcs_xcorr_lags`, csrc/bsseval.cu):
    source x source (Gram matrix blocks), source x estimate (right-hand sides), estimate energy;
  * this module assembles the (nsrc*flen)^2 block-Toeplitz Gram matrix, factors it ONCE (and each diagonal
    block once, for the single-source projections) with a float64 Cholesky on the GPU the lags came from
    (torch.linalg -> cuSOLVER: a library call, this is evaluation tooling, not the separation hot path), solves
    for all estimates together and forms the ratios on the host.

No filtering pass is needed: P_j (projection on source j's delays) and P_all (on all sources') are
orthogonal projections with span_j inside span_all, so with c = G^-1 D
    |s_true + e_spat|^2 = |P_j se|^2 = c_j.D_j        |e_interf|^2 = |P_all se|^2 - |P_j se|^2
    |e_artif|^2 = |se|^2 - |P_all se|^2                |e_interf + e_artif|^2 = |se|^2 - |P_j se|^2
which are the energies of bss_source_crit (:189-199).  `bss_eval_images` / `bss_eval_windowed` are the
multichannel, windowed variant inlined in evaluation/DSD100_eval_only.m:225-306 (30 s windows every
15 s, estimate j against source j, regressors = every channel of every source, SDR / ISR / SIR / SAR),
built from the same device lags with the same identities (plus <P_j se, s_true> = <se, s_true>, s_true
lying in span_j).  The GPU path is mandatory: without libdcs /
a CUDA device the import of the engine raises (no CPU fallback here; the numpy restatement lives in
oracle/bsseval.py as test infrastructure)."""
import ctypes as C
import itertools
import numpy as np

from . import _lib
from .engine import Context, _ptr, _stream_ptr

FLEN = 512


def xcorr_lags(ctx, pairs, L, flen=FLEN, stream=None):
    """pairs: list of (a, b) torch float32 CUDA vectors of L samples -> float64 [npairs, 2*flen-1],
    out[p, li] = sum_t a[t + li - (flen-1)] * b[t]"""
    n = len(pairs)
    pa = (C.c_void_p * n)(*[_ptr(a) for a, _ in pairs])
    pb = (C.c_void_p * n)(*[_ptr(b) for _, b in pairs])
    out = np.empty((n, 2 * flen - 1), dtype=np.float64)
    _lib.check(ctx.lib.dcs_xcorr_lags(ctx.handle, pa, pb, n, int(L), int(flen), out.ctypes.data, _stream_ptr(stream)))
    return out


def _solve(G, D):
    try:
        return np.linalg.solve(G, D)
    except np.linalg.LinAlgError:       # silent / duplicated source (MATLAB warns and carries on)
        return np.linalg.lstsq(G, D, rcond=None)[0]


def _quad_forms(G, D, blocks, device=None):
    """D' G^-1 D for every column of D [n, m], for the whole Gram matrix and for each diagonal block (list of
    slices): ONE factorisation per matrix, shared by all right-hand sides (the reference re-solves per estimate,
    bss_eval_sources.m:147-159).  device = a torch CUDA device: float64 Cholesky + triangular solves on the GPU
    (torch.linalg -> cuSOLVER; the (nsrc*512)^2 systems took 13x the time of the device lag kernel on the host,
    profiles/r2/bsseval_split_host_solve.json); None: numpy on the host (CPU tests).  A singular Gram matrix
    (silent or duplicated source; MATLAB warns and carries on) falls back to least squares on the host."""
    D = np.asarray(D, dtype=np.float64).reshape(G.shape[0], -1)
    out = []
    if device is not None:
        import torch
        Gt = torch.as_tensor(G, dtype=torch.float64, device=device)
        Dt = torch.as_tensor(D, dtype=torch.float64, device=device)
        for sl in [slice(0, G.shape[0])] + list(blocks):
            Gs, Ds = Gt[sl, sl], Dt[sl]
            L, info = torch.linalg.cholesky_ex(Gs)
            if int(info) == 0:
                y = torch.linalg.solve_triangular(L, Ds, upper=False)
                out.append((y * y).sum(dim=0).cpu().numpy())
            else:
                c = np.linalg.lstsq(G[sl, sl], D[sl], rcond=None)[0]
                out.append(np.einsum("ij,ij->j", c, D[sl]))
        return out[0], out[1:]
    for sl in [slice(0, G.shape[0])] + list(blocks):
        c = _solve(G[sl, sl], D[sl])
        out.append(np.einsum("ij,ij->j", c, D[sl]))
    return out[0], out[1:]


def pair_list(n):
    """the signal pairs whose lags the metric needs, as (kind, i, j) with kind 'ss' (true sources i >= j),
    'se' (true source i, estimate j) or 'ee' (estimate i with itself), and the index of each"""
    pairs = [("ss", k1, k2) for k1 in range(n) for k2 in range(k1 + 1)]
    pairs += [("se", k, i) for k in range(n) for i in range(n)]
    pairs += [("ee", i, i) for i in range(n)]
    return pairs, {p if p[0] != "ee" else ("ee", p[1]): q for q, p in enumerate(pairs)}


def ratios_from_lags(R, idx, n, flen, device=None):
    """host part: Gram matrix, solves, energy ratios and the best ordering from the lag table R
    [npairs, 2*flen-1] of `pair_list(n)` -> (SDR, SIR, SAR, perm)"""
    # Gram matrix: block (k1, k2), entry (a, b) = sum_t s_k1[t-a] s_k2[t-b] = R_ss[k1,k2][(b-a) + flen-1]
    lag = (np.arange(flen)[None, :] - np.arange(flen)[:, None]) + flen - 1
    G = np.empty((n * flen, n * flen))
    for k1 in range(n):
        for k2 in range(k1 + 1):
            blk = R[idx["ss", k1, k2]][lag]
            G[k1 * flen:(k1 + 1) * flen, k2 * flen:(k2 + 1) * flen] = blk
            G[k2 * flen:(k2 + 1) * flen, k1 * flen:(k1 + 1) * flen] = blk.T
    SDR, SIR, SAR = (np.zeros((n, n)) for _ in range(3))
    zero = np.float64(0.0)
    # D[:, jest][k*flen + a] = sum_t s_k[t-a] se_jest[t] = R_se[k,jest][flen-1-a]
    Dm = np.stack([np.concatenate([R[idx["se", k, jest]][flen - 1::-1] for k in range(n)]) for jest in range(n)], axis=1)
    blocks = [slice(j * flen, (j + 1) * flen) for j in range(n)]
    p_all_v, p_j_m = _quad_forms(G, Dm, blocks, device)
    with np.errstate(divide="ignore", invalid="ignore"):
        for jest in range(n):
            e_se = np.float64(R[idx["ee", jest]][flen - 1])
            p_all = np.float64(p_all_v[jest])
            for jtrue in range(n):
                p_j = np.float64(p_j_m[jtrue][jest])
                SDR[jest, jtrue] = 10 * np.log10(p_j / np.maximum(e_se - p_j, zero))
                SIR[jest, jtrue] = 10 * np.log10(p_j / np.maximum(p_all - p_j, zero))
                SAR[jest, jtrue] = 10 * np.log10(p_all / np.maximum(e_se - p_all, zero))
    # MATLAB: perm = perms(1:nsrc); [~, popt] = max(meanSIR) (bss_eval_sources.m:56-63).  perms() enumerates in
    # reverse lexicographic order and max() returns the FIRST maximum -- also when every mean is NaN / -Inf
    # (a silent reference or an all-zero estimate), where it returns index 1 instead of failing.
    cands = sorted(itertools.permutations(range(n)), reverse=True)
    best, perm = -np.inf, cands[0]
    for p in cands:
        m = np.mean([SIR[p[j], j] for j in range(n)])
        if m > best:
            best, perm = m, p
    perm = np.array(perm)
    pick = lambda M: np.array([M[perm[j], j] for j in range(n)])
    return pick(SDR), pick(SIR), pick(SAR), perm


def bss_eval_sources(est, ref, flen=FLEN, ctx=None, stream=None):
    """est, ref: [nsrc, L] float32 CUDA tensors (estimated / true sources) ->
    (SDR, SIR, SAR, perm) float64 / int arrays of nsrc entries; estimate perm[j] is matched to true
    source j, the ordering with the best mean SIR (bss_eval_sources.m:54-64)."""
    import torch
    if est.shape != ref.shape or est.dim() != 2:
        raise ValueError("estimated and true sources must both be [nsrc, nsampl]")
    if not (est.is_cuda and ref.is_cuda and est.dtype == torch.float32 and ref.dtype == torch.float32):
        raise ValueError("bss_eval_sources takes float32 CUDA tensors")
    est, ref = est.contiguous(), ref.contiguous()
    n, L = ref.shape
    if ctx is None:
        ctx = Context(ref.device.index or 0)
    kinds, idx = pair_list(n)
    sig = {"ss": (ref, ref), "se": (ref, est), "ee": (est, est)}
    pairs = [(sig[k][0][i], sig[k][1][j]) for k, i, j in kinds]
    R = xcorr_lags(ctx, pairs, L, flen, stream)
    return ratios_from_lags(R, idx, n, flen, device=ref.device)


# ------------------------------------------------------------------------ multichannel images, windowed
def image_pair_list(nsrc, nchan):
    """signal pairs for the images variant: regressors r = j * nchan + c (source j, channel c), estimates
    e = j * nchan + i; kinds 'ss' (r1 >= r2), 'se' (r, e), 'ee' (e, e)"""
    K = nsrc * nchan
    pairs = [("ss", a, b) for a in range(K) for b in range(a + 1)]
    pairs += [("se", r, e) for r in range(K) for e in range(K)]
    pairs += [("ee", e, e) for e in range(K)]
    return pairs, {p if p[0] != "ee" else ("ee", p[1]): q for q, p in enumerate(pairs)}


def images_from_lags(R, idx, nsrc, nchan, flen, device=None):
    """host part of bss_eval_images (DSD100_eval_only.m:240-306) from the lag table of `image_pair_list`
    -> (SDR, ISR, SIR, SAR), each [nsrc]"""
    K = nsrc * nchan
    lag = (np.arange(flen)[None, :] - np.arange(flen)[:, None]) + flen - 1
    G = np.empty((K * flen, K * flen))
    for a in range(K):
        for b in range(a + 1):
            blk = R[idx["ss", a, b]][lag]
            G[a * flen:(a + 1) * flen, b * flen:(b + 1) * flen] = blk
            G[b * flen:(b + 1) * flen, a * flen:(a + 1) * flen] = blk.T
    out = np.zeros((4, nsrc))
    zero = np.float64(0.0)
    # one right-hand side per estimate image e = (source j, channel i); one factorisation of G and of each source's block
    Dm = np.stack([np.concatenate([R[idx["se", r, e]][flen - 1::-1] for r in range(K)]) for e in range(K)], axis=1)
    blocks = [slice(j * nchan * flen, (j + 1) * nchan * flen) for j in range(nsrc)]
    p_all_v, p_j_m = _quad_forms(G, Dm, blocks, device)
    with np.errstate(divide="ignore", invalid="ignore"):
        for j in range(nsrc):
            e_se = e_true = cross = p_j = p_all = zero
            for i in range(nchan):
                e = j * nchan + i
                e_se = e_se + R[idx["ee", e]][flen - 1]
                e_true = e_true + R[idx["ss", e, e]][flen - 1]              # |s_j,i|^2
                cross = cross + Dm[e * flen, e]                             # <se_i, s_j,i>  (delay 0)
                p_all = p_all + p_all_v[e]
                p_j = p_j + p_j_m[j][e]
            out[0, j] = 10 * np.log10(e_true / np.maximum(e_se - 2 * cross + e_true, zero))
            out[1, j] = 10 * np.log10(e_true / np.maximum(p_j - 2 * cross + e_true, zero))
            out[2, j] = 10 * np.log10(p_j / np.maximum(p_all - p_j, zero))
            out[3, j] = 10 * np.log10(p_all / np.maximum(e_se - p_all, zero))
    return out[0], out[1], out[2], out[3]


def window_starts(nsampl, win, ove):
    """first sample of each window: nwin = floor((nsampl - win + 1 + ove) / ove) (DSD100_eval_only.m:228)"""
    nwin = int(np.floor((nsampl - win + 1 + ove) / float(ove)))
    return [k * ove for k in range(max(nwin, 0))]


def bss_eval_images(est, ref, flen=FLEN, ctx=None, stream=None):
    """est, ref: [nsrc, nchan, L] float32 CUDA tensors -> (SDR, ISR, SIR, SAR) of estimate j against source j"""
    import torch
    if est.shape != ref.shape or est.dim() != 3:
        raise ValueError("estimated and true images must both be [nsrc, nchan, nsampl]")
    if not (est.is_cuda and ref.is_cuda and est.dtype == torch.float32 and ref.dtype == torch.float32):
        raise ValueError("bss_eval_images takes float32 CUDA tensors")
    if est.stride(2) != 1 or ref.stride(2) != 1:
        est, ref = est.contiguous(), ref.contiguous()
    nsrc, nchan, L = ref.shape
    if ctx is None:
        ctx = Context(ref.device.index or 0)
    kinds, idx = image_pair_list(nsrc, nchan)
    sig = {"ss": (ref, ref), "se": (ref, est), "ee": (est, est)}
    pairs = [(sig[k][0][a // nchan, a % nchan], sig[k][1][b // nchan, b % nchan]) for k, a, b in kinds]
    return images_from_lags(xcorr_lags(ctx, pairs, L, flen, stream), idx, nsrc, nchan, flen, device=ref.device)


def bss_eval_windowed(est, ref, win, ove, flen=FLEN, ctx=None, stream=None):
    """`bss_eval(ie, i, win, ove)` of DSD100_eval_only.m:225-239 (the caller uses 30 s / 15 s):
    [nsrc, nchan, L] float32 CUDA tensors -> four float64 arrays [nsrc, nwin]"""
    starts = window_starts(ref.shape[-1], win, ove)
    out = np.zeros((4, ref.shape[0], len(starts)))
    if ctx is None and starts:
        ctx = Context(ref.device.index or 0)
    for k, a in enumerate(starts):
        r = bss_eval_images(est[:, :, a:a + win], ref[:, :, a:a + win], flen, ctx, stream)
        for q in range(4):
            out[q, :, k] = r[q]
    return out[0], out[1], out[2], out[3]
