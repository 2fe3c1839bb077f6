"""Drop-in for examples/hiphopss/separate_hhds.py.  The reference file is the DSD100 script byte for byte (one blank
line differs), so this module is that drop-in under the other name."""
import sys

from ..dsd100.separate_dsd import *          # noqa: F401,F403
from ..dsd100.separate_dsd import main, train_auto, build_ca, compute_file, compute_inverse  # noqa: F401

if __name__ == "__main__":
    main(sys.argv[1:])
