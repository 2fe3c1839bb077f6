"""Oracle (TEST INFRASTRUCTURE ONLY): float64 numpy restatement of the reference STFT / iSTFT.

Follows /root/reference/transform.py:
  sinebell        transform.py:35-49
  stft_norm       transform.py:277-335   (= examples/dsd100/separate_dsd.py:49-75)
  istft_norm      transform.py:337-396   (= examples/dsd100/separate_dsd.py:78-111)
  compute_file    transform.py:224-252   (= separate_dsd.py:24-33)
  compute_inverse transform.py:254-274   (= separate_dsd.py:36-41)
The per-frame Python loops are kept on purpose: this module is also the timed CPU baseline
("what the reference does").  Pinned by tests/golden/dsp_*.npz (see oracle/__init__.py).
"""
import numpy as np


def sinebell(lengthWindow):
    # transform.py:48
    return np.sin((np.pi * (np.arange(lengthWindow))) / (1.0 * lengthWindow))


def hanning(n):
    """np.hanning -- the default window of transformFFT (transform.py:221) and of the
    stand-alone scripts (separate_dsd.py:24)."""
    return np.hanning(n)


def blackmanharris(n):
    """scipy.signal.blackmanharris (symmetric) as imported by separate_bach10.py:4; the
    name moved to scipy.signal.windows in current scipy."""
    from scipy.signal import windows
    return windows.blackmanharris(n)


def num_frames(length_data, hopsize):
    # transform.py:309
    return int(np.ceil(length_data / np.double(hopsize)) + 2)


def stft_norm(data, window, hopsize=256.0, nfft=2048.0, fs=44100.0):
    # transform.py:303-335
    lengthWindow = window.size
    lengthData = data.size
    numberFrames = int(np.ceil(lengthData / np.double(hopsize)) + 2)
    newLengthData = int((numberFrames - 1) * hopsize + lengthWindow)
    data = np.concatenate((np.zeros(int(lengthWindow / 2.0)), data))
    data = np.concatenate((data, np.zeros(newLengthData - data.size)))
    numberFrequencies = int(nfft / 2 + 1)
    STFT = np.zeros([numberFrequencies, numberFrames], dtype=complex)
    for n in np.arange(numberFrames):
        beginFrame = int(n * hopsize)
        endFrame = beginFrame + lengthWindow
        frameToProcess = window * data[beginFrame:endFrame]
        STFT[:, n] = np.fft.rfft(frameToProcess, np.int32(nfft))
    return STFT.T


def istft_norm(X, window, analysisWindow=None, hopsize=256.0, nfft=2048.0):
    # transform.py:367-396
    X = X.T
    if analysisWindow is None:
        analysisWindow = window
    lengthWindow = np.array(window.size)
    numberFrequencies, numberFrames = X.shape
    lengthData = int(hopsize * (numberFrames - 1) + lengthWindow)
    normalisationSeq = np.zeros(lengthData)
    data = np.zeros(lengthData)
    for n in np.arange(numberFrames):
        beginFrame = int(n * hopsize)
        endFrame = beginFrame + lengthWindow
        frameTMP = np.fft.irfft(X[:, n], np.int32(nfft))
        frameTMP = frameTMP[:lengthWindow]
        normalisationSeq[beginFrame:endFrame] = (
            normalisationSeq[beginFrame:endFrame] + window * analysisWindow)
        data[beginFrame:endFrame] = data[beginFrame:endFrame] + window * frameTMP
    data = data[int(lengthWindow / 2.0):]
    normalisationSeq = normalisationSeq[int(lengthWindow / 2.0):]
    normalisationSeq[normalisationSeq == 0] = 1.
    data = data / normalisationSeq
    return data


def compute_file(audio, phase=False, frameSize=1024, hopSize=512, sampleRate=44100,
                 window=np.hanning):
    # separate_dsd.py:24-33 / transform.py:243-252 (window may be a callable or an array)
    win = window(frameSize) if callable(window) else window
    X = stft_norm(audio, window=win, hopsize=float(hopSize), nfft=float(frameSize),
                  fs=float(sampleRate))
    mag = np.abs(X)
    mag = mag / np.sqrt(frameSize)
    if phase:
        ph = np.angle(X)
        return mag, ph
    return mag


def compute_inverse(mag, phase, frameSize=1024, hopSize=512, sampleRate=44100,
                    window=np.hanning):
    # separate_dsd.py:36-41 / transform.py:271-274
    win = window(frameSize) if callable(window) else window
    mag = mag * np.sqrt(frameSize)
    Xback = mag * np.exp(1j * phase)
    return istft_norm(Xback, window=win, analysisWindow=win, hopsize=float(hopSize),
                      nfft=float(frameSize))
