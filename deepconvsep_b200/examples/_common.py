"""Shared driver of the stand-alone separation scripts: `train_auto` of
examples/dsd100/separate_dsd.py:239-313 (and its iKala / Bach10 siblings) on the CUDA pipeline."""
import os
import numpy as np
import scipy.io.wavfile

from ..engine import Separator, get_window
from ..models import load_model, FAMILY_DEFAULTS

_cache = {}


def get_separator(model, arch, frame_size, hop, window, scale_factor, time_context, overlap, feat_size):
    key = (os.path.abspath(model), os.path.getmtime(model), arch, frame_size, hop, str(window), scale_factor,
           time_context, overlap)
    if key not in _cache:
        _cache.clear()   # one resident model at a time (Bach10 weights are 856 MB)
        _cache[key] = Separator(load_model(model), arch=arch, frame_size=frame_size, hop=hop, window=window,
                                scale_factor=scale_factor, time_context=time_context, overlap=overlap,
                                patcher="standalone", feat_size=feat_size)
    return _cache[key]


def decode(audioObj, family):
    """scipy.io.wavfile array -> mono float in the reference's (quirky) normalisation:
    divide by iinfo.max -- or by finfo.max for float wavs, which makes those silent
    (separate_dsd.py:277-287) -- then (L+R)/2, or L+R for iKala."""
    if np.issubdtype(audioObj.dtype, np.floating):
        maxv = np.finfo(audioObj.dtype).max
    else:
        maxv = np.iinfo(audioObj.dtype).max
    a = audioObj.astype('float') / maxv
    if family == "ikala":
        return a[:, 0] + a[:, 1]                 # separate_ikala.py:229 (needs a stereo file)
    if a.ndim > 1 and a.shape[1] > 1:
        return (a[:, 0] + a[:, 1]) / 2
    return a if a.ndim == 1 else a[:, 0]


def run(family, filein, outdir, model, scale_factor, time_context, overlap, batch_size, input_size, frame_size, hop,
        out_name):
    """wav in -> one int16 wav per source in `outdir`.  `batch_size` is accepted for signature
    compatibility; the CUDA path has no patch batches."""
    d = FAMILY_DEFAULTS[family]
    sampleRate, audioObj = scipy.io.wavfile.read(filein)
    if sampleRate != 44100:
        print("Sample rate is not 44100")        # separate_dsd.py:313
        return None
    arch = None if family in ("ikala",) else family
    sep = get_separator(model, arch, frame_size, hop, d["window"], scale_factor, time_context, overlap, input_size)
    if audioObj.dtype == np.int16 and family != "ikala":
        stems16 = sep.separate_pcm16(audioObj, downmix=1)          # decode/downmix/encode on the GPU
    else:
        audio = decode(audioObj, family)
        stems = sep.separate(audio)
        stems16 = (stems.astype(np.float64) * np.iinfo(np.int16).max).astype('int16')
    _, filename = os.path.split(filein)
    paths = []
    for i, name in enumerate(sep.sources):
        path = os.path.join(outdir, out_name(filename, name))
        scipy.io.wavfile.write(filename=path, rate=sampleRate, data=stems16[i])
        paths.append(path)
    return paths
