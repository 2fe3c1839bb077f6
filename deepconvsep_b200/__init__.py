"""deepconvsep_b200 -- B200 (sm_100a) native separation hot path of MTG/DeepConvSep.

STFT -> convolutional encoder/decoder -> soft ratio mask -> iSTFT/overlap-add as hand-written
CUDA kernels behind a C ABI (libdcs.so, include/dcs.h), with Python entry points that keep the
reference's names: `transform.transformFFT`, `examples.dsd100.separate_dsd.train_auto/main`, ...
There is no CPU path in this package; it needs the built library and a CUDA device.
"""
from .models import load_model, save_model, infer_arch  # noqa: F401

__all__ = ["load_model", "save_model", "infer_arch", "Separator", "Stft", "Context", "Model"]


def __getattr__(name):
    if name in ("Separator", "Stft", "Context", "Model"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
