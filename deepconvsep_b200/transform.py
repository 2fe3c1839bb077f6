"""Drop-in for the reference's `transform.py` (transform.py:35-396): same names, arguments and
return conventions -- float64 numpy arrays in and out -- with the framed STFT / iSTFT running as
CUDA kernels on the B200 (libdcs: dcs_stft_forward_polar / dcs_istft / dcs_istft_polar).

    from deepconvsep_b200.transform import transformFFT
    tt = transformFFT(frameSize=2048, hopSize=512, sampleRate=44100)
    mag, ph = tt.compute_file(audio, phase=True)
    audio2 = tt.compute_inverse(mag, ph)

Differences from the reference, all deliberate: arithmetic is float32 on the device (results
are returned as float64; ~1e-7 relative error vs the float64 numpy original), frame sizes are
limited to powers of two in [256, 4096], and there is no CPU path.
"""
import re
import numpy as np

from . import engine

_ctx = {}
_plans = {}


def _context(device=None):
    """One libdcs context per device; None = torch's current device (under torchrun each rank has
    called torch.cuda.set_device(local_rank): the plan, the tensors and the stream must all live
    on THAT device)."""
    if device is None:
        import torch
        device = torch.cuda.current_device()
    if device not in _ctx:
        _ctx[device] = engine.Context(device)
    return _ctx[device]


def _plan(window, hop, nfft, syn_window=None, device=None):
    if device is None:
        import torch
        device = torch.cuda.current_device()
    w = np.ascontiguousarray(window, dtype=np.float64)
    n = int(nfft)
    if w.size != n:
        raise ValueError("window length %d != nfft %d: the CUDA STFT needs window.size == nfft" % (w.size, n))
    key = (device, n, int(hop), w.tobytes(), None if syn_window is None else np.asarray(syn_window).tobytes())
    if key not in _plans:
        if len(_plans) > 16:
            _plans.clear()
        _plans[key] = engine.Stft(_context(device), n, int(hop), w, syn_window)
    return _plans[key]


def sinebell(lengthWindow):
    """window(t) = sin(pi t / L), t = 0..L-1   (transform.py:35-49)"""
    return np.sin(np.pi * np.arange(lengthWindow) / float(lengthWindow))


def stft_norm(data, window=None, hopsize=256.0, nfft=2048.0, fs=44100.0):
    """STFT of a 1-D signal, frames centred on sample 0, ceil(L/hop)+2 frames -> complex128
    [frames, nfft/2+1]   (transform.py:277-335)"""
    import torch
    if window is None:
        window = sinebell(2048)
    st = _plan(window, hopsize, nfft)
    x = torch.as_tensor(np.ascontiguousarray(data, dtype=np.float32), device=st.dev)
    X, _ = st.forward(x, want_mag=False)
    return X[:, :st.F].cpu().numpy().astype(np.complex128)


def istft_norm(X, window=None, analysisWindow=None, hopsize=256.0, nfft=2048.0):
    """Overlap-add inverse of stft_norm; the first half window is removed   (transform.py:337-396)"""
    import torch
    if window is None:
        window = sinebell(2048)
    # `window` is the synthesis window, `analysisWindow` the one used by the STFT (transform.py:343-350)
    ana = window if analysisWindow is None else analysisWindow
    st = _plan(ana, hopsize, nfft, syn_window=None if analysisWindow is None else window)
    Xc = np.ascontiguousarray(X, dtype=np.complex64)
    T, F = Xc.shape
    if F != st.F:
        raise ValueError("spectrogram has %d bins, nfft=%d needs %d" % (F, int(nfft), st.F))
    S = torch.zeros((1, T, st.ldf), dtype=torch.complex64, device=st.dev)
    S[0, :, :F] = torch.as_tensor(Xc, device=st.dev)
    return st.inverse(S)[0].cpu().numpy().astype(np.float64)


class Transforms(object):
    """Base class of the feature transforms (transform.py:52-198): holds the analysis settings,
    loops `compute_file` over the columns of an audio matrix and reads / writes the raw float64
    `.data` + `#a\\tb\\tc` `.shape` tensor dumps the training code consumes."""

    def __init__(self, ttype='fft', bins=48, frameSize=1024, hopSize=256, tffmin=25, tffmax=18000, iscale='lin',
                 suffix='', sampleRate=44100, window=np.hanning):
        self.bins = bins
        self.frameSize = frameSize
        self.hopSize = hopSize
        self.fmin = tffmin
        self.fmax = tffmax
        self.iscale = iscale
        self.suffix = suffix
        self.sampleRate = sampleRate
        self.ttype = ttype
        self.window = window(self.frameSize)

    def compute_transform(self, audio, out_path=None, phase=False, save=True):
        """audio[t, i] -> mags[i, T, F] (and phs); saved as <out>_<suffix>_m_.data/.shape (and _p_)
        when `save` and `out_path` are given, returned otherwise   (transform.py:80-131)"""
        self.out_path = out_path
        mags, phs = [], []
        for i in range(audio.shape[1]):
            r = self.compute_file(audio[:, i], phase=phase, sampleRate=self.sampleRate)
            if phase:
                mags.append(r[0])
                phs.append(r[1])
            else:
                mags.append(r)
        mags = np.stack(mags)
        if phase:
            phs = np.stack(phs)
        if save and self.out_path is not None:
            self.saveTensor(mags, '_' + self.suffix + '_m_')
            if phase:
                self.saveTensor(phs, '_' + self.suffix + '_p_')
            return None
        return (mags, phs) if phase else mags

    def compute_file(self, audio, phase=False):
        return None

    def compute_inverse(self, mag, phase):
        return None

    def saveTensor(self, t, name='_cqt_m_'):
        """raw float64 dump + shape file   (transform.py:159-166)"""
        np.asarray(t, dtype=np.float64).tofile(self.out_path.replace('.data', name + '.data'))
        self.shape = tuple(t.shape)
        self.save_shape(self.out_path.replace('.data', name + '.shape'), t.shape)

    def loadTensor(self, name='_cqt_m_'):
        f_in = np.fromfile(self.out_path.replace('.data', name + '.data'))
        shape = self.get_shape(self.out_path.replace('.data', name + '.shape'))
        return f_in.reshape(shape)

    def save_shape(self, shape_file, shape):
        with open(shape_file, 'w') as fout:
            fout.write(u'#' + '\t'.join(str(e) for e in shape) + '\n')

    def get_shape(self, shape_file):
        with open(shape_file, 'rb') as f:
            line = f.readline().decode('ascii')
        if not line.startswith('#'):
            raise IOError('Failed to find shape in file')
        return tuple(map(int, re.findall(r'(\d+)', line)))


class transformFFT(Transforms):
    """STFT features   (transform.py:201-274)"""

    def __init__(self, ttype='fft', bins=48, frameSize=1024, hopSize=256, tffmin=25, tffmax=18000, iscale='lin',
                 suffix='', sampleRate=44100, window=np.hanning):
        super(transformFFT, self).__init__(ttype='fft', bins=bins, frameSize=frameSize, hopSize=hopSize,
                                           tffmin=tffmin, tffmax=tffmax, iscale=iscale, suffix=suffix,
                                           sampleRate=sampleRate, window=window)

    def compute_file(self, audio, phase=False, sampleRate=44100):
        """mag = |STFT| / sqrt(frameSize) [, ph = angle(STFT)]   (transform.py:224-252)"""
        import torch
        st = _plan(self.window, self.hopSize, self.frameSize)
        x = torch.as_tensor(np.ascontiguousarray(audio, dtype=np.float32), device=st.dev)
        if phase:
            mag, ph = st.forward_polar(x)
            return (mag[:, :st.F].cpu().numpy().astype(np.float64), ph[:, :st.F].cpu().numpy().astype(np.float64))
        _, mag = st.forward(x, want_X=False)
        return mag[:, :st.F].cpu().numpy().astype(np.float64)

    def compute_inverse(self, mag, phase, sampleRate=44100):
        """istft_norm(mag * sqrt(frameSize) * exp(j phase))   (transform.py:254-274)"""
        import torch
        st = _plan(self.window, self.hopSize, self.frameSize)
        T, F = mag.shape
        m = torch.zeros((T, st.ldf), dtype=torch.float32, device=st.dev)
        p = torch.zeros((T, st.ldf), dtype=torch.float32, device=st.dev)
        m[:, :F] = torch.as_tensor(np.ascontiguousarray(mag, dtype=np.float32), device=st.dev)
        p[:, :F] = torch.as_tensor(np.ascontiguousarray(phase, dtype=np.float32), device=st.dev)
        return st.inverse_polar(m, p).cpu().numpy().astype(np.float64)
