"""CPU oracle for the DeepConvSep separation hot path -- TEST INFRASTRUCTURE ONLY.

This package is a float64 numpy restatement of the reference algorithm
(MTG/DeepConvSep, `transform.py`, `util.py`, `examples/*/separate_*.py`).  It is the
checker for the CUDA path; it is never the thing shipped or measured as the product.
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  Nothing under `deepconvsep_b200/` imports it.

Parity pinning status
---------------------
* DSP + patcher + overlap-add (`oracle.dsp`, `oracle.patch`): PINNED.  The golden
  vectors in `tests/golden/dsp_*.npz` were produced by executing the *reference's own
  function bodies* (`stft_norm`, `istft_norm`, `compute_file`, `compute_inverse`,
  `generate_overlapadd`, `overlapadd_multi`, `overlapadd`; extracted as source text from
  `/root/reference/transform.py`, `/root/reference/util.py`,
  `/root/reference/examples/dsd100/separate_dsd.py` and
  `/root/reference/examples/ikala/separate_ikala.py` by `tests/golden/make_golden.py`,
  the functions are valid Python 3 although the files are Python 2) -- the oracle must
  reproduce them bit-for-bit / to 1e-15.
* Network arithmetic (`oracle.nets`): **parity unpinned**.  It lives in Theano 0.9 +
  Lasagne (git master), neither of which is vendored in the reference nor installable
  here (Python 3.12, no network), and the reference ships no tests, golden vectors or
  trained weights.  The restatement follows Lasagne's documented layer semantics
   and is cross-checked in `tests/test_oracle_nets.py` against an
  independent torch-autograd formulation (InverseLayer == gradient wrt the layer input);
  `tests/test_oracle_layer_pins.py` holds each layer to a third-party implementation of the
  published definition it stands for (scipy.signal.convolve2d / correlate2d, the adjoint
  identity, torch max_pool2d / max_unpool2d, the DSD net rebuilt from torch library layers).
  That narrows what "unpinned" leaves open to Lasagne deviating from its own documentation.
"""
from . import dsp, patch, nets, pipeline  # noqa: F401
