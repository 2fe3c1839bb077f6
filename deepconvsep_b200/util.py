"""Drop-in for the hot subset of the reference's `util.py`: wav IO (util.py:40-58) and the
patch generator / cross-fade pair used by the trainers' separation branch (util.py:220-327).

In the fused CUDA pipeline (`engine.Separator`, libdcs `dcs_separate_*`) these steps never
materialise: patches are strided *views* of the per-frame encoder activations and the
cross-fade runs inside the mask kernel.  The functions below exist for callers that use them
on their own, with the reference's shapes and semantics, as vectorised numpy (host glue)."""
import numpy as np
import scipy.io.wavfile


def infoAudioScipy(filein):
    sampleRate, audioObj = scipy.io.wavfile.read(filein)
    return len(audioObj), sampleRate, audioObj.dtype


def readAudioScipy(filein):
    """-> (float audio scaled by the dtype's max, sampleRate, dtype)   (util.py:47-54)"""
    sampleRate, audioObj = scipy.io.wavfile.read(filein)
    bitrate = audioObj.dtype
    maxv = np.finfo(bitrate).max if np.issubdtype(bitrate, np.floating) else np.iinfo(bitrate).max
    return audioObj.astype('float') / maxv, sampleRate, bitrate


def writeAudioScipy(fileout, audio_out, sampleRate, bitrate="int16"):
    """(audio * iinfo(bitrate).max).astype(bitrate), no clipping   (util.py:56-58)"""
    maxn = np.iinfo(bitrate).max
    scipy.io.wavfile.write(filename=fileout, rate=sampleRate, data=(audio_out * maxn).astype(bitrate))


def _starts(T, time_context, overlap, limit):
    step = time_context - overlap
    if step <= 0:
        raise ValueError("overlap must be smaller than time_context")
    n = 0 if T <= limit else (T - limit - 1) // step + 1
    return np.arange(n) * step


def generate_overlapadd(allmix, input_size=513, time_context=30, overlap=10, batch_size=32, sampleRate=44100):
    """[T, F] or [C, T, F] -> (fbatch [nbatches, batch_size, C, time_context, F], nchunks); a patch
    starts every time_context-overlap frames while start+overlap < T, zero padded   (util.py:220-248)"""
    allmix = np.asarray(allmix)
    assert input_size == allmix.shape[-1], "Feature size must be the same as the last dimension of the spectrogram"
    x = allmix if allmix.ndim > 2 else allmix[None]
    C, T, F = x.shape
    starts = _starts(T, time_context, overlap, overlap)
    n = len(starts)
    fbatch = np.zeros([int(np.ceil(float(n) / batch_size)), batch_size, C, time_context, F])
    if n:
        pad = np.zeros((C, starts[-1] + time_context, F))
        pad[:, :T] = x
        idx = starts[:, None] + np.arange(time_context)[None, :]
        fbatch.reshape(-1, C, time_context, F)[:n] = pad[:, idx].transpose(1, 0, 2, 3)
    return fbatch, n


def generate_overlapadd_standalone(allmix, input_size=513, time_context=30, overlap=10, batch_size=32, sampleRate=44100):
    """The stand-alone scripts' variant: while start+time_context < T, tail dropped
    (examples/dsd100/separate_dsd.py:114-135; the np.empty tail is zero here)."""
    allmix = np.asarray(allmix)
    T, F = allmix.shape
    starts = _starts(T, time_context, overlap, time_context)
    n = len(starts)
    fbatch = np.zeros([int(np.ceil(float(n) / batch_size)), batch_size, 1, time_context, F])
    if n:
        idx = starts[:, None] + np.arange(time_context)[None, :]
        fbatch.reshape(-1, 1, time_context, F)[:n, 0] = allmix[idx]
    return fbatch, n


def overlapadd_multi(fbatch, obatch, nchunks, overlap=10):
    """fbatch [nbatches, nsources, batch_size, 1, time_context, F] -> sep [nsources,
    nchunks*(time_context-overlap)+time_context, F]   (util.py:297-327)"""
    fbatch = np.asarray(fbatch)
    nsources, F, tc = fbatch.shape[1], fbatch.shape[-1], fbatch.shape[-2]
    step = tc - overlap
    patches = fbatch[:, :, :, 0].transpose(1, 0, 2, 3, 4).reshape(nsources, -1, tc, F)[:, :nchunks]
    sep = np.zeros((nsources, nchunks * step + tc, F))
    # sequential blend, vectorised over sources and bins (each step touches one patch)
    up = np.linspace(0., 1.0, num=overlap)[:, None] if overlap > 0 else np.zeros((0, 1))
    down = up[::-1]
    for k in range(nchunks):
        s = k * step
        if k == 0:
            sep[:, :tc] = patches[:, 0]
        else:
            sep[:, s + overlap:s + tc] = patches[:, k, overlap:]
            sep[:, s:s + overlap] = down * sep[:, s:s + overlap] + up * patches[:, k, :overlap]
    return sep


def overlapadd(fbatch, obatch, nchunks, overlap=10):
    """two-source variant -> (sep1, sep2)   (util.py:251-294)"""
    sep = overlapadd_multi(np.asarray(fbatch)[:, :2], obatch, nchunks, overlap=overlap)
    return sep[0], sep[1]
