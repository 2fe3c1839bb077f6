"""Drop-in for examples/hiphopss/separate_hhds.py: HHDS (hip-hop) four-stem separation; the reference file is the DSD100 script verbatim.

    python -m deepconvsep_b200.examples.dsd100.separate_hhds -i <inputfile> -o <outputdir> -m <path_to_model.pkl>

Same functions and signatures as the reference script; the work happens in the CUDA pipeline
(STFT -> encoder/decoder -> soft mask + cross-fade -> iSTFT), the host only reads and writes wavs."""
import sys
import getopt
import numpy as np

from ...models import load_model                       # noqa: F401  (separate_hhds.py:17-21)
from ...transform import sinebell, stft_norm, istft_norm, transformFFT  # noqa: F401
from ...util import overlapadd_multi                   # noqa: F401  (separate_hhds.py:139-169)
from ...util import generate_overlapadd_standalone as generate_overlapadd  # noqa: F401  (separate_hhds.py:114-135)
from .. import _common

FAMILY = "dsd"
USAGE = 'python separate_hhds.py -i <inputfile> -o <outputdir> -m <path_to_model.pkl>'


def compute_file(audio, phase=False, frameSize=1024, hopSize=512, sampleRate=44100, window=np.hanning):
    """separate_hhds.py:24-33"""
    return transformFFT(frameSize=frameSize, hopSize=hopSize, sampleRate=sampleRate, window=window).compute_file(
        audio, phase=phase)


def compute_inverse(mag, phase, frameSize=1024, hopSize=512, sampleRate=44100, window=np.hanning):
    """separate_hhds.py:36-41"""
    return transformFFT(frameSize=frameSize, hopSize=hopSize, sampleRate=sampleRate, window=window).compute_inverse(
        mag, phase)


def build_ca(input_var=None, batch_size=32, time_context=30, feat_size=513):
    """separate_hhds.py:172-236 built a Lasagne graph; here the network is a fixed CUDA pipeline, so
    this returns the architecture descriptor that dcs_model_create consumes."""
    return {"arch": FAMILY, "time_context": time_context, "feat_size": feat_size, "nsources": 4,
            "layers": ["conv1 50x(1,F)", "conv2 50x(T/2,1)", "dense 128", "3 x dense 800 + InverseLayers", "bias+ReLU"]}


def train_auto(filein, outdir, model, scale_factor=0.3, time_context=30, overlap=20, batch_size=32, input_size=513):
    """separate_hhds.py:239-313: writes vocals.wav, bass.wav, drums.wav, other.wav into outdir."""
    return _common.run(FAMILY, filein, outdir, model, scale_factor, time_context, overlap, batch_size, input_size,
                       frame_size=2 * (input_size - 1), hop=512, out_name=lambda fn, src: src + ".wav")


def main(argv):
    try:
        opts, args = getopt.getopt(argv, "hi:o:m:", ["ifile=", "odir=", "mfile="])
    except getopt.GetoptError:
        print(USAGE)
        sys.exit(2)
    inputfile = outdir = model = None
    for opt, arg in opts:
        if opt == '-h':
            print(USAGE)
            sys.exit()
        elif opt in ("-i", "--ifile"):
            inputfile = arg
        elif opt in ("-o", "--odir"):
            outdir = arg
        elif opt in ("-m", "--mfile"):
            model = arg
    if inputfile is None or outdir is None or model is None:
        print(USAGE)
        sys.exit(2)
    train_auto(inputfile, outdir, model, 0.3, 30, 25, 32, 513)      # separate_hhds.py:332


if __name__ == "__main__":
    main(sys.argv[1:])
