"""Oracle (TEST INFRASTRUCTURE ONLY): the two patchers and the sequential patch cross-fade.

  generate_overlapadd (stand-alone)  examples/dsd100/separate_dsd.py:114-135
  generate_overlapadd (util)         util.py:220-248
  overlapadd_multi                   examples/dsd100/separate_dsd.py:139-169 (= util.py:297-327)
  overlapadd (2 sources)             examples/ikala/separate_ikala.py:138-169 (= util.py:251-294)
One deliberate deviation: the stand-alone patcher allocates with np.empty
(separate_dsd.py:126); the unused tail of the last batch is garbage that is computed on but
never read.  Here it is np.zeros so the oracle is deterministic.
"""
import numpy as np


def generate_overlapadd(allmix, input_size=513, time_context=30, overlap=10, batch_size=32,
                        sampleRate=44100):
    """Stand-alone variant: `while start + time_context < T` -- drops the tail."""
    if input_size != allmix.shape[-1]:
        raise ValueError("Feature size must be the same as the last dimension of the spectrogram")
    i = 0
    start = 0
    while (start + time_context) < allmix.shape[0]:
        i = i + 1
        start = start - overlap + time_context
    fbatch = np.zeros([int(np.ceil(float(i) / batch_size)), batch_size, 1, time_context, input_size])
    i = 0
    start = 0
    while (start + time_context) < allmix.shape[0]:
        fbatch[int(i / batch_size), int(i % batch_size), :, :, :] = allmix[start:start + time_context, :]
        i = i + 1
        start = start - overlap + time_context
    return fbatch, i


def generate_overlapadd_util(allmix, input_size=513, time_context=30, overlap=10, batch_size=32,
                             sampleRate=44100):
    """util.py variant: `while start + overlap < T`, zero-padded, optional channel axis."""
    if len(allmix.shape) > 2:
        nchannels = allmix.shape[0]
    else:
        nchannels = 1
    assert input_size == allmix.shape[-1], \
        "Feature size must be the same as the last dimension of the spectrogram"
    i = 0
    start = 0
    while (start + overlap) < allmix.shape[-2]:
        i = i + 1
        start = start - overlap + time_context
    fbatch = np.zeros([int(np.ceil(float(i) / batch_size)), batch_size, nchannels, time_context, input_size])
    i = 0
    start = 0
    while (start + overlap) < allmix.shape[-2]:
        fbatchend = np.minimum(time_context, allmix.shape[-2] - start)
        end = np.minimum(start + time_context, allmix.shape[-2])
        if len(allmix.shape) > 2:
            fbatch[int(i / batch_size), int(i % batch_size), :, :fbatchend, :] = allmix[:, start:end, :]
        else:
            fbatch[int(i / batch_size), int(i % batch_size), :, :fbatchend, :] = allmix[start:end, :]
        i = i + 1
        start = start - overlap + time_context
    return fbatch, i


def _fade(overlap, input_size):
    window = np.linspace(0., 1.0, num=overlap)
    window = np.concatenate((window, window[::-1]))
    return np.repeat(np.expand_dims(window, axis=1), input_size, axis=1)


def overlapadd_multi(fbatch, obatch, nchunks, overlap=10):
    """fbatch: [nbatches, nsources, batch_size, 1, time_context, F] (np.array of the list of
    per-batch prediction lists, separate_dsd.py:300)."""
    input_size = fbatch.shape[-1]
    time_context = fbatch.shape[-2]
    batch_size = fbatch.shape[2]
    nsources = fbatch.shape[1]
    window = _fade(overlap, input_size)
    sep = np.zeros((nsources, nchunks * (time_context - overlap) + time_context, input_size))
    for s in range(nsources):
        i = 0
        start = 0
        while i < nchunks:
            fbatch1 = fbatch[:, s, :, :, :]
            source = fbatch1[int(i / batch_size), int(i % batch_size), 0, :, :]
            if start == 0:
                sep[s, 0:time_context] = source
            else:
                sep[s, start + overlap:start + time_context] = source[overlap:time_context]
                sep[s, start:start + overlap] = (window[overlap:] * sep[s, start:start + overlap]
                                                 + window[:overlap] * source[:overlap])
            i = i + 1
            start = start - overlap + time_context
    return sep


def overlapadd(fbatch, obatch, nchunks, overlap=10):
    """Two-source variant (iKala)."""
    sep = overlapadd_multi(fbatch[:, :2], obatch, nchunks, overlap=overlap)
    return sep[0], sep[1]


def num_patches(T, time_context, overlap, variant="standalone"):
    """Patch count of either patcher in closed form (tested against the loops)."""
    step = time_context - overlap
    lim = time_context if variant == "standalone" else overlap
    if T <= lim:
        return 0
    return (T - lim - 1) // step + 1


def crossfade_weights(t, P, time_context, overlap):
    """Closed form of the sequential recurrence above: returns
    [(k, weight)] over the patches that contribute to frame t.  Sums to 1 when non-empty."""
    step = time_context - overlap
    k_hi = min(P - 1, t // step)
    k_lo = max(0, -((-(t - time_context + 1)) // step))
    if k_hi < k_lo:
        return []
    up = np.linspace(0., 1.0, num=overlap)
    down = up[::-1]
    ws = []
    acc = {k_lo: 1.0}
    for k in range(k_lo + 1, k_hi + 1):
        p = t - k * step
        for kk in acc:
            acc[kk] *= down[p]
        acc[k] = up[p]
    for k in sorted(acc):
        ws.append((k, acc[k]))
    return ws
