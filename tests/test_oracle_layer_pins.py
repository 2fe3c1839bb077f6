"""The network oracle is "parity unpinned" (Theano 0.9 / Lasagne are not installable here).  What CAN be
pinned is each layer's semantics against third-party implementations of the published definitions the
Lasagne documentation refers to -- not against a second restatement by the same hand:

  Conv2DLayer(flip_filters=True, pad='valid')  = true 2-D convolution           -> scipy.signal.convolve2d
  stride                                       = subsampling of that result     (theano conv2d `subsample`)
  InverseLayer(conv)  = d<conv(x), g>/dx       = the adjoint of a linear map    -> <conv(x), g> == <x, inv(g)>
                                               = full cross-correlation         -> scipy.signal.correlate2d
  MaxPool2DLayer((1, pw)), ignore_border=True  = floor-mode max pooling         -> torch.nn.functional.max_pool2d
  InverseLayer(pool) = MaxPoolGrad                                              -> torch max_unpool2d (tie-free input);
                       every position equal to the window maximum receives the value (theano pool.py MaxPoolGrad)
  DenseLayer = flatten (c, h, w) row-major, x @ W + b, rectify (num_leading_axes=1)
"""
import numpy as np
import scipy.signal
import torch
import torch.nn.functional as Fn

from oracle import nets


def _true_conv(x, W):
    B, C, H, Wd = x.shape
    Fo, _, kh, kw = W.shape
    out = np.zeros((B, Fo, H - kh + 1, Wd - kw + 1))
    for b in range(B):
        for f in range(Fo):
            for c in range(C):
                out[b, f] += scipy.signal.convolve2d(x[b, c], W[f, c], mode="valid")
    return out


def test_conv_layer_is_true_convolution_then_subsampling():
    rng = np.random.default_rng(0)
    for (C, Fo, kh, kw, H, Wd, stride) in ((1, 3, 1, 7, 4, 23, (1, 1)), (2, 4, 3, 1, 9, 6, (1, 1)), (1, 3, 1, 6, 3, 40, (1, 3)),
                                           (4, 2, 1, 5, 2, 33, (1, 4)), (2, 2, 3, 4, 8, 17, (2, 3))):
        x = rng.standard_normal((2, C, H, Wd))
        W = rng.standard_normal((Fo, C, kh, kw))
        want = _true_conv(x, W)[:, :, ::stride[0], ::stride[1]]
        got = nets.conv2d(x, W, stride=stride)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


def test_inverse_of_conv_is_the_adjoint_and_a_full_correlation():
    rng = np.random.default_rng(1)
    for (C, Fo, kh, kw, H, Wd, stride) in ((1, 3, 1, 7, 4, 23, (1, 1)), (2, 4, 3, 1, 9, 6, (1, 1)), (1, 3, 1, 6, 3, 40, (1, 3)),
                                           (4, 2, 1, 5, 2, 34, (1, 4)), (2, 2, 3, 4, 8, 17, (2, 3))):
        x = rng.standard_normal((2, C, H, Wd))
        W = rng.standard_normal((Fo, C, kh, kw))
        y = nets.conv2d(x, W, stride=stride)
        g = rng.standard_normal(y.shape)
        gx = nets.conv2d_inverse(g, W, x.shape, stride=stride)
        assert gx.shape == x.shape
        # the gradient of <conv(x), g> wrt x of a LINEAR map is defined by this identity
        assert abs((y * g).sum() - (x * gx).sum()) <= 1e-11 * abs((y * g).sum()) + 1e-11
        # and, written out: zero-stuff g to the stride-1 grid, full cross-correlation with the same (tied) filters
        sh, sw = stride
        up = np.zeros(g.shape[:2] + ((g.shape[2] - 1) * sh + 1, (g.shape[3] - 1) * sw + 1))
        up[:, :, ::sh, ::sw] = g
        want = np.zeros_like(x)
        for b in range(x.shape[0]):
            for c in range(C):
                acc = 0
                for f in range(Fo):
                    acc = acc + scipy.signal.correlate2d(up[b, f], W[f, c], mode="full")
                want[b, c, :acc.shape[0], :acc.shape[1]] = acc      # input positions no window covers stay 0
        assert np.abs(gx - want).max() <= 1e-12 * np.abs(want).max()


def test_pool_layer_and_its_inverse():
    rng = np.random.default_rng(2)
    for Wd in (16, 17, 19, 161):                      # ignore_border=True: 17 -> 4 windows, the 17th column is dropped
        x = rng.standard_normal((2, 3, 5, Wd))
        xt = torch.tensor(x)
        want, idx = Fn.max_pool2d(xt, (1, 4), return_indices=True)
        got = nets.maxpool_w(x, 4)
        assert np.array_equal(got, want.numpy())
        g = rng.standard_normal(got.shape)
        want_gx = Fn.max_unpool2d(torch.tensor(g), idx, (1, 4), output_size=x.shape[2:]).numpy()
        assert np.array_equal(nets.maxpool_w_inverse(g, x, 4), want_gx)      # tie-free input: one receiver per window
    # ties: theano's MaxPoolGrad gives the value to EVERY position equal to the maximum (silence -> all four)
    x = np.zeros((1, 1, 1, 9))
    x[0, 0, 0, 4:8] = [1.0, 3.0, 3.0, -1.0]
    g = np.array([[[[2.0, 5.0]]]])
    gx = nets.maxpool_w_inverse(g, x, 4)
    assert gx[0, 0, 0].tolist() == [2.0, 2.0, 2.0, 2.0, 0.0, 5.0, 5.0, 0.0, 0.0]


def test_dense_layer_flattening_and_whole_net_against_library_layers():
    """The DSD net assembled from torch's own layers (conv2d on flipped filters, linear, conv_transpose2d with the tied
    flipped filters) -- library kernels, not autograd of the oracle's own formulation."""
    F, tc = 40, 30
    params = nets.make_synthetic_params("dsd", F, tc=tc, seed=4, dtype=np.float64, out_bias=0.05)
    rng = np.random.default_rng(3)
    x = np.abs(rng.standard_normal((2, 1, tc, F))) * 0.3
    p = [torch.tensor(np.asarray(v, dtype=np.float64)) for v in params]
    xt = torch.tensor(x)
    flip = lambda w: torch.flip(w, dims=(2, 3))
    want_pre = nets.predict(params, x, "dsd", return_pre=True)
    a = nets.ARCHS["dsd"]
    assert len(params) == 15              # get_all_param_values order: conv.W, conv.b, BiasLayer.b, ...
    h1 = Fn.conv2d(xt, flip(p[0])) + p[1][None, :, None, None] + p[2][None, :, None, None]
    h2 = Fn.conv2d(h1, flip(p[3])) + p[4][None, :, None, None] + p[5][None, :, None, None]
    z = torch.relu(h2.reshape(2, -1) @ p[6] + p[7])                                  # flatten (c, h, w)
    decs = []
    for s in range(3):
        r = torch.relu(z @ p[8 + 2 * s] + p[9 + 2 * s]).reshape(h2.shape)
        d2 = Fn.conv_transpose2d(r, flip(p[3]))                                      # InverseLayer(conv2): tied filters
        d1 = Fn.conv_transpose2d(d2, flip(p[0]))                                     # InverseLayer(conv1)
        decs.append(d1)
    merged = torch.cat([decs[i] for i in a["dec_of_out"]], dim=1) + p[-1][None, :, None, None]
    assert np.abs(merged.numpy() - want_pre).max() <= 1e-12 * np.abs(want_pre).max()


def _library_net(params, x, arch):
    """Any of the single-stream nets from torch's library layers: conv2d (flipped filters, stride), max_pool2d with
    indices, linear, conv_transpose2d (tied flipped filters, stride, output_padding up to the input width),
    max_unpool2d.  Pre-rectify output, like nets.predict(..., return_pre=True)."""
    a = nets.ARCHS[arch]
    B, nch, tc, F = x.shape
    d = nets.arch_dims(arch, F, tc)
    p = [torch.tensor(np.asarray(v, dtype=np.float64)) for v in params]
    xt = torch.tensor(np.asarray(x, dtype=np.float64))
    flip = lambda w: torch.flip(w, dims=(2, 3))
    s1 = (d["sh1"], d["sw1"])
    h1 = Fn.conv2d(xt, flip(p[0]), stride=s1) + (p[1] + p[2])[None, :, None, None]
    if a["pool"]:
        hp, idx = Fn.max_pool2d(h1, (1, a["pool"]), return_indices=True)
    else:
        hp = h1
    h2 = Fn.conv2d(hp, flip(p[3])) + (p[4] + p[5])[None, :, None, None]
    z = torch.relu(h2.reshape(B, -1) @ p[6] + p[7])
    decs = []
    for s in range(d["ndec"]):
        r = torch.relu(z @ p[8 + 2 * s] + p[9 + 2 * s]).reshape(h2.shape)
        g = Fn.conv_transpose2d(r, flip(p[3]))
        if a["pool"]:
            g = Fn.max_unpool2d(g, idx, (1, a["pool"]), output_size=h1.shape[2:])
        covered = (h1.shape[3] - 1) * s1[1] + p[0].shape[3]
        out = Fn.conv_transpose2d(g, flip(p[0]), stride=s1)
        assert out.shape[3] == covered <= F
        decs.append(Fn.pad(out, (0, F - covered)))                # bins no stride window covers stay 0
    merged = torch.cat([decs[i] for i in a["dec_of_out"]], dim=1) + p[-1][None, :, None, None]
    return merged.numpy()


def test_strided_and_pooled_nets_against_library_layers():
    rng = np.random.default_rng(5)
    for arch, F in (("ikala", 513), ("ikala_nopool", 257), ("bach10", 130), ("bach10_score", 131)):
        params = nets.make_synthetic_params(arch, F, seed=6, dtype=np.float64, out_bias=0.05)
        nch = nets.ARCHS[arch]["nch"]
        x = np.abs(rng.standard_normal((2, nch, 30, F))) * 0.3        # continuous random input: no ties in any window
        want = nets.predict(params, x, arch, return_pre=True)
        got = _library_net(params, x, arch)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), arch
