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
    best, perm = -np.inf, None
    for p in itertools.permutations(range(n)):
        m = np.mean([SIR[p[j], j] for j in range(n)])
        if m > best:
            best, perm = m, p
    perm = np.array(perm)
    pick = lambda M: np.array([M[perm[j], j] for j in range(n)])
    return pick(SDR), pick(SIR), pick(SAR), perm
