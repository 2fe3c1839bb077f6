"""BSS-Eval 3.0 `bss_eval_sources` for single-channel sources, restated in numpy -- TEST INFRASTRUCTURE
(like the rest of oracle/: only tests/, smoke() and bench.py's CPU legs may import it).

What it restates (evaluation/bss_eval/bss_eval_sources.m, used by evaluation/evaluate_SS_iKala.m:58-59
and evaluation/Bach10_eval_only.m:94; this is what "SDR" means in BASELINE.json's parity bar):

  project             :110-159   least-squares projection of an estimate on the span of the true sources
                                 delayed by 0..flen-1 samples; Gram matrix and right-hand side from FFT
                                 auto-/cross-correlations of the zero-padded signals, one dense solve
  bss_decomp_mtifilt  :70-106    estimate = s_true + e_spat + e_interf + e_artif  (flen = 512)
  bss_source_crit     :163-199   SDR / SIR / SAR energy ratios in dB
  bss_eval_sources    :1-66      all (estimate, source) pairs, then the ordering with the best mean SIR
and the windowed multichannel "images" variant inlined in evaluation/DSD100_eval_only.m:
  bss_eval            :225-239   30 s windows every 15 s (the caller passes 30*fs, 15*fs, :171-177)
  bss_eval_images     :240-255   estimate j against source j (no ordering search), 512-tap projection
  project             :257-295   regressors = every channel of every source, each delayed by 0..511
  bss_image_crit      :297-306   SDR / ISR / SIR / SAR

PARITY UNPINNED against the MATLAB code itself: there is no MATLAB / Octave (and no mir_eval / museval)
in this image.  It is pinned instead by tests/test_oracle_bsseval.py: `project` against a brute-force
time-domain least-squares projection, exact recovery of FIR-filtered sources, the analytic SIR of a
two-source leak, permutation recovery and gain invariance.

Deviation from a literal transcription (same numbers, less work): the reference rebuilds the Gram
matrix for each of the nsrc^2 (estimate, source) pairs; it depends on the true sources only, so it is
built once, and the single-source projections use its diagonal blocks."""
import itertools
import numpy as np

FLEN = 512


def _gram(S, flen):
    """S [n, L] -> (G [n*flen, n*flen], Sf, fftlen): inner products between the delayed, zero-padded rows
    (bss_eval_sources.m:120-136).  Block (k1, k2), entry (a, b) = sum_t s_k1[t-a] * s_k2[t-b]."""
    n, L = S.shape
    fftlen = 1 << int(np.ceil(np.log2(L + flen - 1)))
    Sf = np.fft.rfft(S, fftlen, axis=1)
    lag = (np.arange(flen)[None, :] - np.arange(flen)[:, None]) % fftlen      # [a, b] -> (b - a) mod fftlen
    G = np.empty((n * flen, n * flen))
    for k1 in range(n):
        for k2 in range(k1 + 1):
            r = np.fft.irfft(Sf[k1] * np.conj(Sf[k2]), fftlen)
            blk = r[lag]
            G[k1 * flen:(k1 + 1) * flen, k2 * flen:(k2 + 1) * flen] = blk
            G[k2 * flen:(k2 + 1) * flen, k1 * flen:(k1 + 1) * flen] = blk.T
    return G, Sf, fftlen


def _rhs(se, Sf, fftlen, flen):
    """inner products between the estimate and the delayed sources (bss_eval_sources.m:138-145) -> [n*flen]"""
    sef = np.fft.rfft(se, fftlen)
    idx = (-np.arange(flen)) % fftlen
    return np.concatenate([np.fft.irfft(Sf[k] * np.conj(sef), fftlen)[idx] for k in range(Sf.shape[0])])


def _solve(G, D):
    try:
        return np.linalg.solve(G, D)
    except np.linalg.LinAlgError:       # a silent / duplicated source: MATLAB warns and carries on
        return np.linalg.lstsq(G, D, rcond=None)[0]


def _filter_sum(C, S, flen):
    """sum_k conv(C[k], s_k) over the zero-padded length L + flen - 1 (bss_eval_sources.m:151-157)"""
    n, L = S.shape
    m = L + flen - 1
    fl = 1 << int(np.ceil(np.log2(m + flen - 1)))
    acc = (np.fft.rfft(C, fl, axis=1) * np.fft.rfft(S, fl, axis=1)).sum(axis=0)
    return np.fft.irfft(acc, fl)[:m]


def project(se, S, flen=FLEN):
    """least-squares projection of se [L] on span{ s_k delayed by 0..flen-1 } -> [L + flen - 1]"""
    S = np.atleast_2d(np.asarray(S, dtype=np.float64))
    G, Sf, fftlen = _gram(S, flen)
    C = _solve(G, _rhs(np.asarray(se, dtype=np.float64), Sf, fftlen, flen))
    return _filter_sum(C.reshape(S.shape[0], flen), S, flen)


def source_crit(s_true, e_spat, e_interf, e_artif):
    """(SDR, SIR, SAR) in dB (bss_eval_sources.m:189-199)"""
    s_filt = s_true + e_spat
    with np.errstate(divide="ignore"):
        sdr = 10 * np.log10(np.sum(s_filt ** 2) / np.sum((e_interf + e_artif) ** 2))
        sir = 10 * np.log10(np.sum(s_filt ** 2) / np.sum(e_interf ** 2))
        sar = 10 * np.log10(np.sum((s_filt + e_interf) ** 2) / np.sum(e_artif ** 2))
    return sdr, sir, sar


def bss_eval_sources(se, s, flen=FLEN):
    """se, s: [nsrc, L] estimated / true sources -> (SDR, SIR, SAR, perm), each [nsrc]; estimate perm[j]
    is matched to true source j (the ordering with the best mean SIR, bss_eval_sources.m:54-64)."""
    se = np.asarray(se, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    if se.shape != s.shape or s.ndim != 2:
        raise ValueError("estimated and true sources must both be [nsrc, nsampl]")
    n, L = s.shape
    G, Sf, fftlen = _gram(s, flen)
    pad = np.zeros(flen - 1)
    SDR, SIR, SAR = (np.zeros((n, n)) for _ in range(3))
    for jest in range(n):
        D = _rhs(se[jest], Sf, fftlen, flen)
        p_all = _filter_sum(_solve(G, D).reshape(n, flen), s, flen)
        se_pad = np.concatenate([se[jest], pad])
        for jtrue in range(n):
            blk = slice(jtrue * flen, (jtrue + 1) * flen)
            cj = _solve(G[blk, blk], D[blk])
            p_j = _filter_sum(cj[None, :], s[jtrue:jtrue + 1], flen)
            s_true = np.concatenate([s[jtrue], pad])
            e_spat = p_j - s_true
            e_interf = p_all - s_true - e_spat
            e_artif = se_pad - s_true - e_spat - e_interf
            SDR[jest, jtrue], SIR[jest, jtrue], SAR[jest, jtrue] = source_crit(s_true, e_spat, e_interf, e_artif)
    # perms(1:nsrc) enumerates in reverse lexicographic order, max() keeps the first maximum and returns
    # index 1 when every mean is NaN / -Inf (bss_eval_sources.m:56-63)
    cands = sorted(itertools.permutations(range(n)), reverse=True)
    best, perm = -np.inf, cands[0]
    for p in cands:
        m = np.mean([SIR[p[j], j] for j in range(n)])
        if m > best:
            best, perm = m, p
    perm = np.array(perm)
    pick = lambda M: np.array([M[perm[j], j] for j in range(n)])
    return pick(SDR), pick(SIR), pick(SAR), perm


# ------------------------------------------------------------------------ multichannel images, windowed
def project_images(se, S, flen=FLEN):
    """se [nchan, L]; S [nsrc, nchan, L] -> projection [nchan, L + flen - 1] of each channel of se on the
    span of all channels of all sources in S delayed by 0..flen-1 (DSD100_eval_only.m:257-295; regressor
    order = channel fastest, as MATLAB's reshape of [nsampl, nchan, nsrc])."""
    se = np.atleast_2d(np.asarray(se, dtype=np.float64))
    S = np.asarray(S, dtype=np.float64)
    nsrc, nchan, L = S.shape
    R = S.reshape(nsrc * nchan, L)                       # row j*nchan + c
    G, Rf, fftlen = _gram(R, flen)
    out = np.empty((se.shape[0], L + flen - 1))
    for i in range(se.shape[0]):
        C = _solve(G, _rhs(se[i], Rf, fftlen, flen))
        out[i] = _filter_sum(C.reshape(R.shape[0], flen), R, flen)
    return out


def image_crit(s_true, e_spat, e_interf, e_artif):
    """(SDR, ISR, SIR, SAR) in dB over all channels (DSD100_eval_only.m:297-306)"""
    with np.errstate(divide="ignore"):
        sdr = 10 * np.log10(np.sum(s_true ** 2) / np.sum((e_spat + e_interf + e_artif) ** 2))
        isr = 10 * np.log10(np.sum(s_true ** 2) / np.sum(e_spat ** 2))
        sir = 10 * np.log10(np.sum((s_true + e_spat) ** 2) / np.sum(e_interf ** 2))
        sar = 10 * np.log10(np.sum((s_true + e_spat + e_interf) ** 2) / np.sum(e_artif ** 2))
    return sdr, isr, sir, sar


def bss_eval_images(ie, i, flen=FLEN):
    """ie, i: [nsrc, nchan, L] estimated / true source images -> (SDR, ISR, SIR, SAR), each [nsrc];
    estimate j is scored against source j (DSD100_eval_only.m:240-255)."""
    ie = np.asarray(ie, dtype=np.float64)
    i = np.asarray(i, dtype=np.float64)
    if ie.shape != i.shape or i.ndim != 3:
        raise ValueError("estimated and true images must both be [nsrc, nchan, nsampl]")
    nsrc, nchan, L = i.shape
    pad = np.zeros((nchan, flen - 1))
    out = np.zeros((4, nsrc))
    for j in range(nsrc):
        s_true = np.concatenate([i[j], pad], axis=1)
        e_spat = project_images(ie[j], i[j:j + 1], flen) - s_true
        e_interf = project_images(ie[j], i, flen) - s_true - e_spat
        e_artif = np.concatenate([ie[j], pad], axis=1) - s_true - e_spat - e_interf
        out[:, j] = image_crit(s_true, e_spat, e_interf, e_artif)
    return out[0], out[1], out[2], out[3]


def window_starts(nsampl, win, ove):
    """first sample of each evaluation window (DSD100_eval_only.m:228,236: nwin = floor((nsampl-win+1+ove)/ove))"""
    nwin = int(np.floor((nsampl - win + 1 + ove) / float(ove)))
    return [k * ove for k in range(max(nwin, 0))]


def bss_eval_windowed(ie, i, win, ove, flen=FLEN):
    """`bss_eval(ie, i, win, ove)` of DSD100_eval_only.m:225-239 -> four arrays [nsrc, nwin]"""
    ie = np.asarray(ie, dtype=np.float64)
    i = np.asarray(i, dtype=np.float64)
    starts = window_starts(i.shape[-1], win, ove)
    out = np.zeros((4, i.shape[0], len(starts)))
    for k, a in enumerate(starts):
        r = bss_eval_images(ie[:, :, a:a + win], i[:, :, a:a + win], flen)
        for q in range(4):
            out[q, :, k] = r[q]
    return out[0], out[1], out[2], out[3]
