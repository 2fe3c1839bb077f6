"""Oracle (TEST INFRASTRUCTURE ONLY): the whole `train_auto` separation of the stand-alone
scripts, float64, structured exactly like the reference executes it -- per-frame STFT loop,
patch copy loop, batches of 32 through the network, sequential cross-fade, one iSTFT per
source.  This is both the parity checker for the CUDA pipeline and the timed CPU baseline.

  DSD100   examples/dsd100/separate_dsd.py:239-313   (hanning, N=1024, overlap 25)
  iKala    examples/ikala/separate_ikala.py:194-256  (hanning, N=1024, overlap 20, L+R)
  Bach10   examples/bach10/separate_bach10.py:232-306 (blackmanharris, N=4096, overlap 25)
  util patcher variant: examples/dsd100/trainCNN.py:300-333
  stereo / ILD variant: examples/dsd100_2ch_ILD/trainCNN_ILD_DSD100.py:291-327
"""
import numpy as np
from . import dsp, patch, nets


def decode_wav_array(audioObj, family="dsd"):
    """int/float PCM array as returned by scipy.io.wavfile.read -> mono float64
    (separate_dsd.py:277-287; iKala sums L+R without halving, separate_ikala.py:229)."""
    try:
        maxv = np.finfo(audioObj.dtype).max
    except ValueError:
        maxv = np.iinfo(audioObj.dtype).max
    a = audioObj.astype('float') / maxv
    if family == "ikala":
        return a[:, 0] + a[:, 1]
    if a.ndim > 1 and a.shape[1] > 1:
        return (a[:, 0] + a[:, 1]) / 2
    return a if a.ndim == 1 else a[:, 0]


def separate(audio, params, arch, frameSize=1024, hopSize=512, window=np.hanning,
             scale_factor=0.3, time_context=30, overlap=25, batch_size=32,
             patcher="standalone", return_spec=False, count_kinks=False, pool_bits=None):
    """mono float64 audio [L] -> stems float64 [nsrc, L].
    pool_bits (max-pool net, parity tests only): the device's tie bits uint8 [T, WP, C>=30] (bit r: position
    4*jp+r of the window received the value); adopted in ill-conditioned windows only, see
    nets.maxpool_w_inverse; separate.last_pool_stats holds the counts."""
    a = nets.ARCHS[arch]
    mag, ph = dsp.compute_file(audio, phase=True, frameSize=frameSize, hopSize=hopSize,
                               window=window)
    mag = scale_factor * mag.astype(np.float32)          # separate_dsd.py:290 (float32!)
    gen = patch.generate_overlapadd if patcher == "standalone" else patch.generate_overlapadd_util
    batches, nchunks = gen(mag, input_size=mag.shape[-1], time_context=time_context,
                           overlap=overlap, batch_size=batch_size)
    pool_stats = {}
    if pool_bits is not None:
        # per-frame device decisions -> per-patch [B, C, tc, wp, 4]; frames beyond T (zero padding) and the unused
        # tail of the last batch keep the float64 decision (their input is exactly constant)
        step_, T_ = time_context - overlap, mag.shape[0]
        bits = np.asarray(pool_bits)[:, :, :30]
        dev_frames = np.stack([(bits >> r) & 1 for r in range(4)], axis=-1).astype(bool)       # [T, WP, C, 4]
        dev_frames = dev_frames.transpose(2, 0, 1, 3)                                         # [C, T, WP, 4]

        def dev_hits_of(bi, b):
            C_, T2, WP_, _ = dev_frames.shape
            out = np.zeros((b.shape[0], C_, time_context, WP_, 4), dtype=bool)
            valid = np.zeros((b.shape[0], 1, time_context, 1, 1), dtype=bool)
            for i in range(b.shape[0]):
                k = bi * batch_size + i
                if k >= nchunks:
                    break
                t0 = k * step_
                n = max(0, min(time_context, T_ - t0, T2 - t0))
                out[i, :, :n] = dev_frames[:, t0:t0 + n]
                valid[i, 0, :n] = True
            return out, valid
        pres = [nets.predict(params, b, arch, return_pre=True, pool_dev=dev_hits_of(bi, b), pool_stats=pool_stats)
                for bi, b in enumerate(batches)]
    else:
        pres = [nets.predict(params, b, arch, return_pre=True) for b in batches] if count_kinks else None
    separate.last_pool_stats = pool_stats
    output = [nets.predict_function2(params, b, arch, pred=None if pres is None else nets.relu(pres[i]))
              for i, b in enumerate(batches)]
    output = np.array(output)                            # [nb, nsrc, B, 1, tc, F]
    kink_energy = 0.0
    if count_kinks:
        nk, left = 0, nchunks
        step = time_context - overlap
        kmap = np.zeros((max(len(ph), nchunks * step + time_context), mag.shape[-1]), dtype=bool)
        for bi, b in enumerate(batches):
            nb = max(0, min(left, batch_size))
            flag = nets.near_kink(pres[bi][:nb], a["mask"], a["nsrc"])  # [nb, tc, F]
            nk += int(flag.sum())
            kink_energy += float((flag * b[:nb].sum(axis=1) ** 2).sum())
            for i in np.nonzero(flag.reshape(nb, -1).any(axis=1))[0]:
                k0 = (bi * batch_size + int(i)) * step
                kmap[k0:k0 + time_context] |= flag[i]
            left -= batch_size
        separate.last_kinks = nk
        # time-frequency bins some covering patch flags (the blended mask of such a bin is ill-conditioned)
        separate.last_kink_map = kmap[:len(ph)]
    if nchunks == 0:
        mm = np.zeros((a["nsrc"], len(ph), mag.shape[-1]))
    else:
        mm = patch.overlapadd_multi(output, batches, nchunks, overlap=overlap)
    if count_kinks:
        # worst-case relative error a flip of every flagged bin can cause, per stem: the mask of a
        # flagged bin moves by at most 1, so the blended magnitude moves by at most the mixture's
        separate.last_kink_bound = [float(np.sqrt(kink_energy / max(float((mm[i] ** 2).sum()), 1e-300)))
                                    for i in range(a["nsrc"])]
    stems = []
    for i in range(a["nsrc"]):
        m = mm[i, :len(ph)]
        if m.shape[0] < len(ph):                          # cannot happen with the stock patchers
            m = np.concatenate([m, np.zeros((len(ph) - m.shape[0], m.shape[1]))])
        audio_out = dsp.compute_inverse(m / scale_factor, ph, frameSize=frameSize,
                                        hopSize=hopSize, window=window)
        if len(audio_out) > len(audio):
            audio_out = audio_out[:len(audio)]
        stems.append(audio_out)
    stems = np.stack(stems)
    if return_spec:
        return stems, mag, ph, mm
    return stems


def separate_score(audio, filters, params, frameSize=4096, hopSize=512, window=None, scale_factor=0.2,
                   time_context=30, overlap=25, batch_size=32, count_kinks=False, return_spec=False):
    """Score-informed separation branch of examples/bach10_scoreinformed/trainCNNrwc.py:384-416:
    filters [4, T, F] float32 (LargeDatasetMask2.filterSpec) -> input channels filter*mag (float32
    products), util's zero-padded patcher on the 3-D tensor, network + Bach10 mask rule on the sum
    of the channels, cross-fade, inverse STFT with the mixture phase."""
    if window is None:
        window = dsp.blackmanharris
    arch = "bach10_score"
    a = nets.ARCHS[arch]
    mag, ph = dsp.compute_file(audio, phase=True, frameSize=frameSize, hopSize=hopSize, window=window)
    mag = scale_factor * mag.astype(np.float32)
    masks = np.ones((4, mag.shape[0], mag.shape[1]))
    for j in range(4):
        masks[j] = np.asarray(filters[j], dtype=np.float32) * mag
    batches, nchunks = patch.generate_overlapadd_util(masks, input_size=masks.shape[-1], time_context=time_context,
                                                      overlap=overlap, batch_size=batch_size)
    pres = [nets.predict(params, b, arch, return_pre=True) for b in batches] if count_kinks else None
    output = np.array([nets.predict_function2(params, b, arch, pred=None if pres is None else nets.relu(pres[i]))
                       for i, b in enumerate(batches)])
    kink_energy = 0.0
    if count_kinks:
        nk, left = 0, nchunks
        step = time_context - overlap
        kmap = np.zeros((max(len(ph), nchunks * step + time_context), mag.shape[-1]), dtype=bool)
        for bi, b in enumerate(batches):
            nb = max(0, min(left, batch_size))
            flag = nets.near_kink(pres[bi][:nb], a["mask"], a["nsrc"])
            nk += int(flag.sum())
            kink_energy += float((flag * b[:nb].sum(axis=1) ** 2).sum())
            for i in np.nonzero(flag.reshape(nb, -1).any(axis=1))[0]:
                k0 = (bi * batch_size + int(i)) * step
                kmap[k0:k0 + time_context] |= flag[i]
            left -= batch_size
        separate_score.last_kinks = nk
        separate_score.last_kink_map = kmap[:len(ph)]
    mm = patch.overlapadd_multi(output, batches, nchunks, overlap=overlap)
    if count_kinks:
        separate_score.last_kink_bound = [float(np.sqrt(kink_energy / max(float((mm[i] ** 2).sum()), 1e-300))) for i in range(4)]
    stems = []
    for i in range(4):
        audio_out = dsp.compute_inverse(mm[i, :len(ph)] / scale_factor, ph, frameSize=frameSize, hopSize=hopSize, window=window)
        stems.append(audio_out[:len(audio)] if len(audio_out) > len(audio) else audio_out)
    if return_spec:
        return np.stack(stems), mag, ph, mm
    return np.stack(stems)


def separate_stereo(audio, params, frameSize=1024, hopSize=512, window=np.hanning, scale_factor=0.3,
                    time_context=30, overlap=25, batch_size=32, count_kinks=False, return_spec=False):
    """Separation loop of the stereo / ILD trainer (trainCNN_ILD_DSD100.py:299-327):
    audio float64 [L, 2] -> stems float64 [L, nsrc, 2] (`sep_audio`).  One STFT per channel
    (`compute_transform`), util's zero-padded patcher on the [2, T, F] tensor, one network pass per
    batch, then per channel the 4-source cross-fade and one iSTFT per (source, channel) with that
    channel's mixture phase.  return_spec: also (mag [nch,T,F], phases [nch][T,F], blended magnitudes
    [nch][nsrc,T',F])."""
    a = nets.ARCHS["dsd_ild"]
    nch = audio.shape[1]
    mags, phs = [], []
    for j in range(nch):                                   # transform.py:105-119
        m, p = dsp.compute_file(audio[:, j], phase=True, frameSize=frameSize, hopSize=hopSize, window=window)
        mags.append(m)
        phs.append(p)
    mag = scale_factor * np.stack(mags).astype(np.float32)  # :304
    batches, nchunks = patch.generate_overlapadd_util(mag, input_size=mag.shape[-1], time_context=time_context,
                                                      overlap=overlap, batch_size=batch_size)
    pres = [nets.predict(params, b, "dsd_ild", return_pre=True) for b in batches] if count_kinks else None
    output = np.array([nets.predict_function_ild(params, b, pred=None if pres is None else nets.relu(pres[i]))
                       for i, b in enumerate(batches)])   # [nb, nch, B, nsrc, tc, F]
    kink_energy = np.zeros(nch)
    if count_kinks:     # bins whose mask sits on its discontinuity (all outputs of a channel vanish), see separate()
        nk, left = 0, nchunks
        step = time_context - overlap
        T = phs[0].shape[0]
        kmap = np.zeros((nch, max(T, nchunks * step + time_context), mag.shape[-1]), dtype=bool)
        for bi, b in enumerate(batches):
            nb = max(0, min(left, batch_size))
            pre = pres[bi][:nb]
            for j in range(nch):
                flag = nets.near_kink(pre[:, j::nch], a["mask"], a["nsrc"])
                nk += int(flag.sum())
                kink_energy[j] += float((flag * b[:nb, j] ** 2).sum())
                for i in np.nonzero(flag.reshape(nb, -1).any(axis=1))[0]:
                    k0 = (bi * batch_size + int(i)) * step
                    kmap[j, k0:k0 + time_context] |= flag[i]
            left -= batch_size
        separate_stereo.last_kinks = nk
        separate_stereo.last_kink_bound = np.zeros((a["nsrc"], nch))
        separate_stereo.last_kink_map = kmap[:, :T]
    sep = np.zeros((audio.shape[0], a["nsrc"], nch))
    mms = []
    for j in range(nch):
        mm = patch.overlapadd_multi(np.swapaxes(output[:, j:j + 1], 1, 3), batches, nchunks, overlap=overlap)
        mms.append(mm)
        if count_kinks:
            for i in range(a["nsrc"]):
                separate_stereo.last_kink_bound[i, j] = float(np.sqrt(kink_energy[j] / max(float((mm[i] ** 2).sum()), 1e-300)))
        for i in range(a["nsrc"]):
            audio_out = dsp.compute_inverse(mm[i, :phs[j].shape[0]] / scale_factor, phs[j], frameSize=frameSize,
                                            hopSize=hopSize, window=window)
            sep[:, i, j] = audio_out[:audio.shape[0]]
    if return_spec:
        return sep, mag, phs, mms
    return sep


def synth_mixture(seconds, seed, sr=44100):
    """Seeded synthetic 4-stem mixture: harmonic tone with vibrato,
    low sine bursts, noise bursts, pink-ish noise; int16-quantised like a wav file.
    Returns (mixture float64 [L], stems float64 [4, L])."""
    rng = np.random.default_rng(seed)
    L = int(round(seconds * sr))
    t = np.arange(L) / sr
    f0 = rng.uniform(110, 440)
    vib = 1 + 0.01 * np.sin(2 * np.pi * 5 * t)
    s1 = sum(np.sin(2 * np.pi * f0 * h * t * vib + rng.uniform(0, 6.28)) / h for h in range(1, 11))
    fb = rng.uniform(55, 110)
    s2 = np.sin(2 * np.pi * fb * t) * (np.sin(2 * np.pi * rng.uniform(1, 2) * t) > 0)
    rate = rng.uniform(2, 4)
    env = np.exp(-((t * rate) % 1.0) / (rate * 0.010))
    s3 = rng.standard_normal(L) * env
    w = rng.standard_normal(L)
    s4 = np.cumsum(w) * 0.02
    s4 = s4 - np.convolve(s4, np.ones(512) / 512, mode="same") + 0.3 * w
    stems = np.stack([s1, s2, s3, s4])
    stems *= 0.2 / np.maximum(np.abs(stems).max(axis=1, keepdims=True), 1e-12)
    mix = stems.sum(axis=0)
    mix = np.round(mix * 32767).astype(np.int16).astype('float') / 32767
    return mix, stems
