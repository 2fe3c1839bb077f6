"""Host side of the B200 separation path: thin Python over the C ABI (libdcs.so).

`Stft` mirrors transform.transformFFT's numerical core, `Separator` mirrors `train_auto` of the
stand-alone scripts (examples/dsd100/separate_dsd.py:239-313).  torch is used only as a
device-memory / stream container; numpy arrays go through the *_host entry points."""
import ctypes as C
import numpy as np

from . import _lib
from .models import infer_arch, FAMILY_DEFAULTS


def get_window(window, n):
    """np.hanning / scipy blackmanharris (symmetric) / sinebell by name, callable or array."""
    if isinstance(window, str):
        if window in ("hanning", "hann"):
            return np.hanning(n)
        if window == "blackmanharris":
            from scipy.signal import windows
            return windows.blackmanharris(n)
        if window == "sinebell":
            return np.sin((np.pi * (np.arange(n))) / (1.0 * n))
        raise ValueError("unknown window %r" % window)
    if callable(window):
        return np.asarray(window(n), dtype=np.float64)
    w = np.asarray(window, dtype=np.float64)
    if w.size != n:
        raise ValueError("window has %d samples, frame size is %d" % (w.size, n))
    return w


def _ptr(a):
    """Raw address of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def _stream_ptr(stream=None, device=None):
    """cudaStream_t of `stream`, or of torch's current stream ON `device` (the ctx's device: the
    current stream of another device would be a handle the library cannot launch on)."""
    import torch
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return C.c_void_p(s.cuda_stream)


class Context(object):
    """One dcs_ctx = one device + the workspace of one in-flight pipeline."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.dcs_create(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)

    def launch_count(self):
        return int(self.lib.dcs_launch_count(self.handle))

    def workspace_bytes(self):
        return int(self.lib.dcs_workspace_bytes(self.handle))

    def profile(self, enable):
        _lib.check(self.lib.dcs_profile(self.handle, 1 if enable else 0))

    def profile_read(self, max_n=4096):
        """[(stage name, milliseconds)] recorded since profiling was enabled (synchronises)."""
        names = C.create_string_buffer(64 * max_n)
        ms = (C.c_float * max_n)()
        n = self.lib.dcs_profile_read(self.handle, names, len(names), ms, max_n)
        if n < 0:
            _lib.check(n)
        nm = names.value.decode().split("\n")[:n]
        return list(zip(nm, [float(ms[i]) for i in range(n)]))

    def gemm(self, A, B, bias=None, relu=False, engine=1, stream=None):
        """torch float32 cuda A [M,K] (row stride A.stride(0)) x numpy B [K,N] -> torch [M,N]."""
        import torch
        M, K = A.shape
        Bh = np.ascontiguousarray(B, dtype=np.float32)
        N = Bh.shape[1]
        bh = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        Cd = torch.empty((M, N), dtype=torch.float32, device=A.device)
        _lib.check(self.lib.dcs_gemm_f32(self.handle, int(engine), _ptr(A), A.stride(0), Bh.ctypes.data, N,
                                         None if bh is None else bh.ctypes.data, _ptr(Cd), N, M, N, K, int(relu),
                                         _stream_ptr(stream, self.device)))
        return Cd

    def close(self):
        if getattr(self, "handle", None):
            self.lib.dcs_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Stft(object):
    """STFT plan (frame size, hop, analysis/synthesis windows)."""

    def __init__(self, ctx, frame_size, hop, window=np.hanning, syn_window=None):
        self.ctx, self.lib = ctx, ctx.lib
        self.N, self.hop = int(frame_size), int(hop)
        self.device = ctx.device
        self.F = self.N // 2 + 1
        self.ldf = int(self.lib.dcs_padded_bins(self.N))
        self.window = np.ascontiguousarray(get_window(window, self.N), dtype=np.float64)
        syn = None if syn_window is None else np.ascontiguousarray(get_window(syn_window, self.N), dtype=np.float64)
        h = C.c_void_p()
        _lib.check(self.lib.dcs_stft_plan(ctx.handle, self.N, self.hop, self.window.ctypes.data,
                                          None if syn is None else syn.ctypes.data, C.byref(h)))
        self.handle = h

    @property
    def dev(self):
        import torch
        return torch.device("cuda", self.device)

    def num_frames(self, L):
        return int(self.lib.dcs_num_frames(int(L), self.hop))

    def out_length(self, T):
        return (T - 1) * self.hop + self.N - self.N // 2

    # ---- device-tensor API (torch) ----
    def forward(self, audio, mag_scale=1.0, want_X=True, want_mag=True, stream=None):
        """audio: torch float32 cuda [L] -> (X complex64 [T, ldf] | None, mag float32 [T, ldf] | None)"""
        import torch
        L = audio.numel()
        T = self.num_frames(L)
        X = torch.empty((T, self.ldf), dtype=torch.complex64, device=audio.device) if want_X else None
        mag = torch.empty((T, self.ldf), dtype=torch.float32, device=audio.device) if want_mag else None
        _lib.check(self.lib.dcs_stft_forward(self.handle, _ptr(audio), L, _ptr(X), _ptr(mag), float(mag_scale),
                                             self.ldf, _stream_ptr(stream, self.device)))
        return X, mag

    def forward_polar(self, audio, mag_scale=1.0, stream=None):
        import torch
        L = audio.numel()
        T = self.num_frames(L)
        mag = torch.empty((T, self.ldf), dtype=torch.float32, device=audio.device)
        ph = torch.empty((T, self.ldf), dtype=torch.float32, device=audio.device)
        _lib.check(self.lib.dcs_stft_forward_polar(self.handle, _ptr(audio), L, _ptr(mag), _ptr(ph), float(mag_scale),
                                                   self.ldf, _stream_ptr(stream, self.device)))
        return mag, ph

    def inverse(self, S, num_out=None, stream=None):
        """S: torch complex64 cuda [nsrc, T, ldf] (or [T, ldf]) -> float32 [nsrc, num_out]"""
        import torch
        if S.dim() == 2:
            S = S.unsqueeze(0)
        nsrc, T, ldf = S.shape
        assert S.is_contiguous() and ldf >= self.F
        n = self.out_length(T) if num_out is None else int(num_out)
        out = torch.empty((nsrc, n), dtype=torch.float32, device=S.device)
        _lib.check(self.lib.dcs_istft(self.handle, _ptr(S), nsrc, T, ldf, T * ldf, _ptr(out), n, n, _stream_ptr(stream, self.device)))
        return out

    def inverse_polar(self, mag, phase, mag_scale=1.0, num_out=None, stream=None):
        import torch
        T, ldf = mag.shape
        n = self.out_length(T) if num_out is None else int(num_out)
        out = torch.empty((n,), dtype=torch.float32, device=mag.device)
        _lib.check(self.lib.dcs_istft_polar(self.handle, self.ctx.handle, _ptr(mag), _ptr(phase), float(mag_scale), T, ldf,
                                            _ptr(out), n, _stream_ptr(stream, self.device)))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.lib.dcs_stft_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Model(object):
    """Device-resident network built from a Lasagne parameter list."""

    def __init__(self, ctx, params, arch=None, feat_size=None, time_context=None):
        self.ctx, self.lib = ctx, ctx.lib
        try:
            ia, iF, itc = infer_arch(params, feat_size)   # time_context None: the one the weights were trained with
        except ValueError:
            if arch is None or feat_size is None:
                raise
            ia, iF, itc = arch, feat_size, 30
        a = ia if arch is None else arch
        F = feat_size if (arch is not None and feat_size is not None) else iF
        self.arch, self.F, self.tc = a, int(F), int(time_context or itc)
        arrs = [np.ascontiguousarray(p, dtype=np.float32) for p in params]
        n = len(arrs)
        ptrs = (C.c_void_p * n)(*[x.ctypes.data for x in arrs])
        shapes = np.ones((n, 4), dtype=np.int64)
        ndims = np.zeros(n, dtype=np.int32)
        for i, x in enumerate(arrs):
            ndims[i] = x.ndim
            shapes[i, :x.ndim] = x.shape
        h = C.c_void_p()
        _lib.check(self.lib.dcs_model_create(ctx.handle, _lib.ARCH_IDS[a], self.F, self.tc, n, ptrs,
                                             shapes.ctypes.data, ndims.ctypes.data, C.byref(h)))
        self.handle = h
        self.nsrc = int(self.lib.dcs_model_nsources(h))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.dcs_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Separator(object):
    """train_auto() as an object: build once (weights uploaded, plan made), call many times."""

    def __init__(self, params, arch=None, frame_size=None, hop=None, window=None, scale_factor=0.3,
                 time_context=None, overlap=None, patcher="standalone", device=0, feat_size=None):
        self.ctx = Context(device)
        if arch is None and frame_size is not None and feat_size is None:
            feat_size = frame_size // 2 + 1
        self.model = Model(self.ctx, params, arch=arch, feat_size=feat_size, time_context=time_context)
        d = FAMILY_DEFAULTS[self.model.arch]
        self.frame_size = int(frame_size or 2 * (self.model.F - 1))
        self.hop = int(hop or d["hopSize"])
        self.window = window if window is not None else d["window"]
        self.overlap = int(d["overlap"] if overlap is None else overlap)
        self.scale_factor = float(scale_factor)
        self.patcher = _lib.PATCHER_IDS[patcher]
        self.stft = Stft(self.ctx, self.frame_size, self.hop, self.window)
        self.nsrc = self.model.nsrc
        self.sources = d["sources"]
        self.lib = self.ctx.lib

    # ---- host buffers (numpy): H2D + pipeline + D2H inside the call ----
    def separate(self, audio, out=None):
        """audio: 1-D float array (any float dtype) -> float32 [nsrc, L].  `audio` / `out` may be
        pinned (torch.from_numpy(...).pin_memory() views) for asynchronous copies."""
        a = np.ascontiguousarray(audio, dtype=np.float32)
        L = a.size
        if out is None:
            out = np.empty((self.nsrc, L), dtype=np.float32)
        assert out.dtype == np.float32 and out.shape == (self.nsrc, L) and out.flags.c_contiguous
        _lib.check(self.lib.dcs_separate_host(self.ctx.handle, self.model.handle, self.stft.handle, a.ctypes.data, L,
                                              self.scale_factor, self.overlap, self.patcher, out.ctypes.data, L,
                                              _stream_ptr(None, self.ctx.device)))
        return out

    def separate_pcm16(self, pcm, downmix=1, out=None):
        """int16 wav samples [L] or [L, channels] -> int16 [nsrc, L] (train_auto's wav contract)."""
        p = np.ascontiguousarray(pcm, dtype=np.int16)
        L = p.shape[0]
        ch = 1 if p.ndim == 1 else p.shape[1]
        if out is None:
            out = np.empty((self.nsrc, L), dtype=np.int16)
        _lib.check(self.lib.dcs_separate_pcm16_host(self.ctx.handle, self.model.handle, self.stft.handle, p.ctypes.data, L,
                                                    ch, int(downmix if ch > 1 else 0), self.scale_factor, self.overlap,
                                                    self.patcher, out.ctypes.data, L, _stream_ptr(None, self.ctx.device)))
        return out

    def separate_pcm16_batch(self, clips, downmix=1, outs=None):
        """Several clips through the context's multi-clip scheduler (dcs_separate_batch_pcm16_host): H2D of clip i+1,
        the kernels of clip i and D2H of clip i-1 overlap.  clips: list of int16 arrays [L] or [L, channels] (same
        channel count; pinned for real overlap) -> list of int16 [nsrc, L]."""
        ps = [np.ascontiguousarray(c, dtype=np.int16) for c in clips]
        n = len(ps)
        if n == 0:
            return []
        ch = 1 if ps[0].ndim == 1 else ps[0].shape[1]
        assert all((1 if p_.ndim == 1 else p_.shape[1]) == ch for p_ in ps), "all clips must have the same channel count"
        Ls = np.array([p_.shape[0] for p_ in ps], dtype=np.int64)
        if outs is None:
            outs = [np.empty((self.nsrc, int(L)), dtype=np.int16) for L in Ls]
        assert all(o.dtype == np.int16 and o.shape == (self.nsrc, int(L)) and o.flags.c_contiguous for o, L in zip(outs, Ls))
        pin = (C.c_void_p * n)(*[p_.ctypes.data for p_ in ps])
        pout = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        _lib.check(self.lib.dcs_separate_batch_pcm16_host(self.ctx.handle, self.model.handle, self.stft.handle, n, pin,
                                                          Ls.ctypes.data, ch, int(downmix if ch > 1 else 0), self.scale_factor,
                                                          self.overlap, self.patcher, pout, Ls.ctypes.data,
                                                          _stream_ptr(None, self.ctx.device)))
        return outs

    # ---- device buffers (torch tensors) ----
    def separate_device(self, audio, out=None, stream=None):
        """audio: torch float32 cuda [L] -> torch float32 cuda [nsrc, L]; asynchronous."""
        import torch
        L = audio.numel()
        if out is None:
            out = torch.empty((self.nsrc, L), dtype=torch.float32, device=audio.device)
        _lib.check(self.lib.dcs_separate_audio(self.ctx.handle, self.model.handle, self.stft.handle, _ptr(audio), L,
                                               self.scale_factor, self.overlap, self.patcher, _ptr(out), out.stride(0),
                                               _stream_ptr(stream, self.ctx.device)))
        return out

    def separate_score(self, audio, filters, out=None, stream=None):
        """Score-informed Bach10: audio float [L] (numpy or cuda tensor) + normalised score filters
        [4, T, F] float32 (deepconvsep_b200.score.score_filters; or a cuda tensor [4, T, ldf] already on the
        device) -> stems float32 [4, L] (same kind as `audio`).  The four input channels are formed on the device."""
        import torch
        host = not hasattr(audio, "is_cuda")
        x = torch.as_tensor(np.ascontiguousarray(audio, dtype=np.float32), device=self.stft.dev) if host else audio
        L = x.numel()
        T = self.stft.num_frames(L)
        if hasattr(filters, "is_cuda"):      # already on the device, padded rows: [4, T, ldf] float32
            fd = filters
            assert fd.is_cuda and fd.dtype == torch.float32 and fd.is_contiguous() and tuple(fd.shape) == (4, T, self.stft.ldf)
        else:
            f = np.asarray(filters, dtype=np.float32)
            assert f.shape == (4, T, self.model.F), (f.shape, (4, T, self.model.F))
            fd = torch.zeros((4, T, self.stft.ldf), dtype=torch.float32, device=x.device)
            fd[:, :, :self.model.F] = torch.as_tensor(f, device=x.device)
        if out is None or host:
            outd = torch.empty((self.nsrc, L), dtype=torch.float32, device=x.device)
        else:
            outd = out
        _lib.check(self.lib.dcs_separate_audio_score(self.ctx.handle, self.model.handle, self.stft.handle, _ptr(x), L,
                                                     _ptr(fd), self.scale_factor, self.overlap, self.patcher, _ptr(outd),
                                                     outd.stride(0), _stream_ptr(stream, self.ctx.device)))
        return outd.cpu().numpy() if host else outd

    def separate_stereo(self, audio, out=None, stream=None):
        """Stereo / ILD network (examples/dsd100_2ch_ILD/trainCNN_ILD_DSD100.py:299-327): audio float
        [L, 2] (numpy) or [2, L] (cuda tensor) -> `sep_audio` float32 [L, nsrc, 2] (numpy) or the device
        planes [nsrc * 2, L] ordered (source, channel) (cuda tensor in -> cuda tensor out)."""
        import torch
        host = not hasattr(audio, "is_cuda")
        if host:
            a = np.asarray(audio, dtype=np.float32)
            assert a.ndim == 2 and a.shape[1] == 2, a.shape
            x = torch.as_tensor(np.ascontiguousarray(a.T), device=self.stft.dev)
        else:
            x = audio.contiguous()
            assert x.dim() == 2 and x.shape[0] == 2 and x.dtype == torch.float32
        L = x.shape[1]
        outd = out if (out is not None and not host) else torch.empty((self.nsrc * 2, L), dtype=torch.float32, device=x.device)
        _lib.check(self.lib.dcs_separate_audio_stereo(self.ctx.handle, self.model.handle, self.stft.handle, _ptr(x), x.stride(0), L,
                                                      self.scale_factor, self.overlap, self.patcher, _ptr(outd), outd.stride(0),
                                                      _stream_ptr(stream, self.ctx.device)))
        if not host:
            return outd
        return np.ascontiguousarray(outd.cpu().numpy().reshape(self.nsrc, 2, L).transpose(2, 0, 1))

    def separate_tapped(self, audio, filters=None, pool=False):
        """Parity-test entry: the same pipeline as separate() / separate_score() / separate_stereo() with
        the spectrum tap on (dcs_set_spectrum_tap) -> (stems as that call returns them, masked spectra
        complex64 numpy [nplanes, T, F] -- the tensors the inverse STFT of THIS call consumed).
        pool=True (max-pool net): also the tie bits uint8 [T, WP, 32] of this call (dcs_set_pool_tap)."""
        import torch
        a = np.asarray(audio)
        L = a.shape[0]
        T = self.stft.num_frames(L)
        nplanes = self.nsrc * (2 if self.model.arch == "dsd_ild" else 1)
        tap = torch.zeros((nplanes, T, self.stft.ldf), dtype=torch.complex64, device=self.stft.dev)
        _lib.check(self.lib.dcs_set_spectrum_tap(self.ctx.handle, _ptr(tap), tap.numel()))
        bits = None
        if pool:
            assert self.model.arch == "ikala", "only the max-pool network has routing decisions to tap"
            WP = ((self.model.F - 30) // 3 + 1) // 4
            bits = torch.zeros((T, WP, 32), dtype=torch.uint8, device=self.stft.dev)
            _lib.check(self.lib.dcs_set_pool_tap(self.ctx.handle, _ptr(bits), bits.numel()))
        try:
            if self.model.arch == "bach10_score":
                out = self.separate_score(a, filters)
            elif self.model.arch == "dsd_ild":
                out = self.separate_stereo(a)
            else:
                out = self.separate(a)
            torch.cuda.synchronize(self.stft.dev)
        finally:
            _lib.check(self.lib.dcs_set_spectrum_tap(self.ctx.handle, None, 0))
            _lib.check(self.lib.dcs_set_pool_tap(self.ctx.handle, None, 0))
        S = tap[:, :, :self.model.F].cpu().numpy()
        return (out, S, bits.cpu().numpy()) if pool else (out, S)

    def separate_spec(self, mag, X, stream=None):
        """scaled magnitude [T, ldf] + mixture STFT [T, ldf] -> masked spectra complex64 [nsrc, T, ldf]"""
        import torch
        T, ldf = mag.shape
        S = torch.empty((self.nsrc, T, ldf), dtype=torch.complex64, device=mag.device)
        _lib.check(self.lib.dcs_separate_spec(self.ctx.handle, self.model.handle, _ptr(mag), _ptr(X), T, ldf, self.overlap,
                                              self.patcher, _ptr(S), T * ldf, _stream_ptr(stream, self.ctx.device)))
        return S

    def num_patches(self, T):
        return int(self.lib.dcs_num_patches(int(T), self.model.tc, self.overlap, self.patcher))
