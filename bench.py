#!/usr/bin/env python
"""bench.py -- audio-seconds separated per second, DSD100 4-stem configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (STFT -> encoder/decoder -> soft mask + cross-fade ->
iSTFT/OLA) over one batch of `--clips` synthetic 180 s mono 44.1 kHz mixtures per GPU
(BASELINE.json configs[1]: frameSize=2048, hop=512, time_context=30, overlap 25, 4 sources).
`value`   : whole-job audio-s/s with the inputs already resident in HBM (CUDA events, max over ranks);
`e2e`     : the same metric through the C-ABI host-buffer call with pinned host memory, H2D and D2H
            inside the timed region: int16 wav samples in / int16 stems out (train_auto's contract,
            dcs_separate_pcm16_host); the float32-buffer call (dcs_separate_host) is reported beside it;
`roofline`: dominant kernel, timed live with CUDA events on its own stream (dcs_profile);
`cpu_baseline`: the float64 numpy oracle (the restated reference path) on a bounded sample.
`--impl reference` times that CPU path on all host cores (the reference itself -- Python 2 +
Theano 0.9 + Lasagne -- cannot be installed here).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 44100
METRIC = "audio_seconds_separated_per_second"
_emit = print
UNIT = "audio-s/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=8, help="clips per step per GPU")
    ap.add_argument("--seconds", type=float, default=180.0)
    ap.add_argument("--frame-size", type=int, default=2048)
    ap.add_argument("--e2e-streams", type=int, default=3)
    ap.add_argument("--device-streams", type=int, default=1,
                    help="contexts / CUDA streams the device-resident loop spreads a step's clips over (kernels of "
                         "different clips overlap: the single-wave mask and iSTFT kernels leave tails)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dsd_param_shapes(F, tc=30):
    h2 = tc - tc // 2 + 1
    return ([(50, 1, 1, F), (50,), (50,), (50, 50, tc // 2, 1), (50,), (50,), (50 * h2, 128), (128,)]
            + [(128, 50 * h2), (50 * h2,)] * 3 + [(4,)])


def synthetic_params(F, seed=0):
    """Random-init DSD100 weights (Lasagne GlorotUniform; output bias scaled so that the masks vary)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    shapes = dsd_param_shapes(F)
    out = []
    for i, s in enumerate(shapes):
        if len(s) == 4:
            a = math.sqrt(6.0 / ((s[0] + s[1]) * s[2] * s[3]))
        elif len(s) == 2:
            a = math.sqrt(6.0 / (s[0] + s[1]))
        else:
            a = 0.002 if i == len(shapes) - 1 else 0.1
        out.append(rng.uniform(-a, a, size=s).astype(np.float32))
    return out


def synth_clip_device(L, seed, device):
    """Synthetic 4-stem mixture generated on the device: harmonic tone with vibrato, gated low
    sine, decaying noise bursts, broadband noise; int16-quantised like a wav file."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.arange(L, device=device, dtype=torch.float32) / SR
    r = torch.rand(8, generator=g, device=device)
    f0 = 110 + 330 * r[0]
    ph = 2 * math.pi * f0 * t * (1 + 0.01 * torch.sin(2 * math.pi * 5 * t))
    s1 = sum(torch.sin(ph * h + 6.28 * r[1] * h) / h for h in range(1, 11))
    fb = 55 + 55 * r[2]
    s2 = torch.sin(2 * math.pi * fb * t) * (torch.sin(2 * math.pi * (1 + r[3]) * t) > 0)
    rate = 2 + 2 * r[4]
    env = torch.exp(-((t * rate) % 1.0) / (rate * 0.010))
    s3 = torch.randn(L, generator=g, device=device) * env
    s4 = torch.randn(L, generator=g, device=device)
    mix = torch.zeros(L, device=device)
    for s in (s1, s2, s3, s4):
        mix += 0.2 * s / s.abs().max().clamp_min(1e-9)
    return torch.round(mix * 32767).clamp(-32768, 32767) / 32767.0


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def stage_work(stage, L, N, hop=512, tc=30, ov=25):
    """Algorithmic (bytes, flops) of one launch of a pipeline stage for one clip."""
    F = N // 2 + 1
    T = -(-L // hop) + 2
    step = tc - ov
    P = (T - tc - 1) // step + 1
    fft = 2.5 * N * math.log2(N)
    w = {
        "stft_fwd": (L * 4 + T * F * 12, T * fft),
        "enc_conv1_gemm": (T * F * 4 + T * 50 * 4 + F * 50 * 4, 2.0 * T * F * 50),
        "enc_conv2_gemm": (T * 50 * 8, 2.0 * (T - 14) * 750 * 50),
        "bottleneck_gemm": (T * 50 * 4 + P * 128 * 4 + 800 * 128 * 4, 2.0 * P * 800 * 128),
        "dec_dense_gemm": (P * 128 * 4 + P * 2400 * 4 + 128 * 2400 * 4, 2.0 * P * 128 * 2400),
        "dec_convT2_gemm": (P * 3 * 44 * 50 * 4 + P * 90 * 50 * 4, 2.0 * P * 90 * 750 * 50),
        "dec_convT1_mask_xfade": (T * F * 8 + P * 90 * 50 * 4 + 4 * T * F * 8, 2.0 * P * 90 * 50 * F),
        "istft_ola": (4 * T * F * 8 + 4 * L * 4, 4 * T * fft),
    }
    return w[stage]


def load_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(job):
    seconds, seed, N = job
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    import numpy as np
    from oracle import nets, pipeline
    params = synthetic_params(N // 2 + 1, 0)
    mix, _ = pipeline.synth_mixture(seconds, seed)
    t = time.perf_counter()
    out = pipeline.separate(mix, params, "dsd", frameSize=N, hopSize=512, window=np.hanning, overlap=25)
    dt = time.perf_counter() - t
    assert out.shape[0] == 4
    return dt


def cpu_reference_run(N, steps, warmup, clip_seconds, procs):
    """The restated reference path on the host cores: `procs` single-threaded workers, one clip
    each per step."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        for w in range(warmup):
            pool.map(_cpu_worker, [(min(clip_seconds, 2.0), 7000 + i, N) for i in range(procs)])
        t0 = time.perf_counter()
        for s in range(steps):
            pool.map(_cpu_worker, [(clip_seconds, 1000 + s * procs + i, N) for i in range(procs)])
        dt = time.perf_counter() - t0
    return steps * procs * clip_seconds / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    clip_s = 10.0
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    value, dt = cpu_reference_run(args.frame_size, steps, warm, clip_s, cores)
    sample = "%d steps x %d workers x one %.0f s clip each (float64 numpy restatement of separate_dsd.py, BLAS 1 thread/worker)" % (
        steps, cores, clip_s)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, cpu=True),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference's own runtime (Python 2.7 + Theano 0.9 + Lasagne) is not installable here; this is "
                "the oracle port of its CPU path",
    }
    _emit(json.dumps(line))


def workload_config(args, cpu=False):
    return {"workload": "BASELINE configs[1]: DSD100 4-source separation, mono 44.1 kHz, frameSize=%d hop=512 "
                        "time_context=30 overlap=25, %s" % (args.frame_size,
                                                            "bounded CPU sample" if cpu else
                                                            "%d clips x %.0f s per step per GPU" % (args.clips, args.seconds)),
            "frame_size": args.frame_size, "hop": 512, "time_context": 30, "overlap": 25, "nsrc": 4,
            "clips_per_step_per_gpu": args.clips, "clip_seconds": args.seconds,
            "l2_policy": "inputs larger than L2 (%.0f MB of audio per step, >0.8 GB of intermediates per clip)" % (
                args.clips * args.seconds * SR * 4 / 1e6),
            "parallelism": "clips sharded over %d GPU(s), no data-path collective" % args.gpus}


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from deepconvsep_b200.engine import Separator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    N, B, K, W = args.frame_size, args.clips, args.steps, max(args.warmup, 3)
    L = int(round(args.seconds * SR))
    params = synthetic_params(N // 2 + 1, 0)
    sep = Separator(params, frame_size=N, hop=512, window="hanning", overlap=25, device=local)
    clips = [synth_clip_device(L, 1000 + rank * B + i, dev) for i in range(B)]
    outs = torch.empty((B, 4, L), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nd = max(1, args.device_streams)
    dworkers = [sep] + [Separator(params, frame_size=N, hop=512, window="hanning", overlap=25, device=local)
                        for _ in range(nd - 1)]
    dstreams = [torch.cuda.Stream(device=dev) for _ in range(nd)] if nd > 1 else []

    def step_device():
        if nd == 1:
            for i in range(B):
                sep.separate_device(clips[i], outs[i])
            return
        cur = torch.cuda.current_stream()
        for s_ in dstreams:
            s_.wait_stream(cur)
        for i in range(B):
            dworkers[i % nd].separate_device(clips[i], outs[i], stream=dstreams[i % nd])
        for s_ in dstreams:
            cur.wait_stream(s_)

    # ---- device-resident throughput -------------------------------------------------------
    for _ in range(W):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    launches0 = sum(w.ctx.launch_count() for w in dworkers)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        step_device()
    e1.record()
    barrier()
    t_host1 = time.perf_counter()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    launches = sum(w.ctx.launch_count() for w in dworkers) - launches0
    clocks = sampler.stop(t_host0, t_host1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    audio_s = world * B * K * args.seconds
    value = audio_s / (ms * 1e-3)
    finite = bool(torch.isfinite(outs).all().item())

    # ---- per-stage timing (CUDA events inside libdcs, on the launching stream) --------------
    sep.ctx.profile(True)
    for _ in range(2):     # one stream: stage times must not include the overlap with other clips
        for i in range(B):
            sep.separate_device(clips[i], outs[i])
    torch.cuda.synchronize()
    recs = sep.ctx.profile_read()
    sep.ctx.profile(False)
    stages = {}
    for name, t in recs:
        stages.setdefault(name, []).append(t)
    stage_ms = {k: sum(v) / len(v) for k, v in stages.items()}
    tot = sum(stage_ms.values())
    dom = max(stage_ms, key=stage_ms.get)
    hbm_peak, tensor_peak, peak_src = measured_peaks()
    by, fl = stage_work(dom, L, N)
    ach = by / (stage_ms[dom] * 1e-3) / 1e9
    traffic = load_traffic().get("%s@N%d" % (dom, N))
    def roof(stage):
        b_, f_ = stage_work(stage, L, N)
        t_ = stage_ms[stage] * 1e-3
        if stage.endswith("_gemm"):   # dense contractions: tensor pipe (3xTF32 = 3 MMAs at half the bf16 rate)
            a_ = f_ / t_ / 1e12
            return {"kernel": stage, "bound": "tensor", "achieved": a_, "peak": tensor_peak, "unit": "TFLOP/s",
                    "frac": a_ / tensor_peak, "note": "fp32-accurate 3xTF32: the ceiling of this arithmetic is peak/6"}
        a_ = b_ / t_ / 1e9
        return {"kernel": stage, "bound": "hbm", "achieved": a_, "peak": hbm_peak, "unit": "GB/s", "frac": a_ / hbm_peak}
    roofline = roof(dom)
    roofline.update({"traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": by,
                     "algorithmic_flops_per_launch": fl, "launch_ms": stage_ms[dom],
                     "share_of_step": stage_ms[dom] / tot,
                     "all_stages": [dict(roof(k), share_of_step=stage_ms[k] / tot) for k in stage_ms]})
    stage_table = {k: {"ms": v, "share": v / tot, "gbps": stage_work(k, L, N)[0] / (v * 1e-3) / 1e9,
                       "tflops": stage_work(k, L, N)[1] / (v * 1e-3) / 1e12} for k, v in stage_ms.items()}

    # ---- end to end through the host-buffer C-ABI call --------------------------------------
    ns = max(1, args.e2e_streams)
    workers = [sep] + [Separator(params, frame_size=N, hop=512, window="hanning", overlap=25, device=local)
                       for _ in range(ns - 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(ns)]
    # (a) the wav contract of train_auto (separate_dsd.py:275-287,307-309): int16 samples in, int16 stems out
    # (b) float32 host buffers (dcs_separate_host)
    h_in16 = [torch.empty(L, dtype=torch.int16).pin_memory() for _ in range(B)]
    h_out16 = [torch.empty((4, L), dtype=torch.int16).pin_memory() for _ in range(B)]
    h_in = [torch.empty(L, dtype=torch.float32).pin_memory() for _ in range(B)]
    h_out = [torch.empty((4, L), dtype=torch.float32).pin_memory() for _ in range(B)]
    for i in range(B):
        h_in[i].copy_(clips[i].cpu())
        h_in16[i].copy_(torch.round(clips[i].cpu() * 32767).to(torch.int16))
    np_in, np_out = [t.numpy() for t in h_in], [t.numpy() for t in h_out]
    np_in16, np_out16 = [t.numpy() for t in h_in16], [t.numpy() for t in h_out16]

    def e2e_worker(w, nsteps, pcm):
        torch.cuda.set_device(local)
        with torch.cuda.stream(streams[w]):
            for _ in range(nsteps):
                for i in range(w, B, ns):
                    if pcm:
                        workers[w].separate_pcm16(np_in16[i], out=np_out16[i])
                    else:
                        workers[w].separate(np_in[i], out=np_out[i])

    def e2e_run(nsteps, pcm):
        ths = [threading.Thread(target=e2e_worker, args=(w, nsteps, pcm)) for w in range(ns)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    def e2e_measure(pcm):
        e2e_run(2, pcm)
        barrier()
        l0 = sum(w.ctx.launch_count() for w in workers)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        e2e_run(K, pcm)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        f1.record()
        barrier()
        nl = sum(w.ctx.launch_count() for w in workers) - l0
        t = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), nl

    ems_f32, _ = e2e_measure(False)
    ems, e2e_launches = e2e_measure(True)
    e2e_value = audio_s / (ems * 1e-3)
    # result checks on the host copies: the four stems add up to the mixture where masks cover it
    chk = float(np.abs(np_out[0][:, 44100:88200].sum(0) - np_in[0][44100:88200]).max())
    chk16 = int(np.abs(np_out16[0][:, 44100:88200].astype(np.int32).sum(0) - np_in16[0][44100:88200]).max())

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded sample --------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        secs = 20.0
        t0 = time.perf_counter()
        dt = _cpu_worker_threads(secs, N)
        cpu = {"value": secs / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "one %.0f s clip, frameSize=%d, float64 numpy oracle (per-frame FFT loops, batches of 32 "
                         "patches, BLAS threads = all cores), %.1f s of CPU time" % (secs, N, time.perf_counter() - t0)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": dict(workload_config(args), device_streams=nd),
            "x_realtime": value, "gpu_launches": int(launches), "outputs_finite": finite,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * L * 2, "d2h_bytes_per_step": B * 4 * L * 2,
                    "ms_per_step": ems / K, "streams": ns, "gpu_launches": int(e2e_launches),
                    "api": "dcs_separate_pcm16_host: pinned int16 wav samples in, int16 stems out (train_auto's wav contract)",
                    "stem_sum_max_abs_err_lsb": chk16,
                    "float32_buffers": {"value": audio_s / (ems_f32 * 1e-3), "ms_per_step": ems_f32 / K,
                                        "h2d_bytes_per_step": B * L * 4, "d2h_bytes_per_step": B * 4 * L * 4,
                                        "api": "dcs_separate_host", "stem_sum_max_abs_err": chk}},
            "roofline": roofline, "stages": stage_table, "cpu_baseline": cpu,
        }
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _cpu_worker_threads(seconds, N):
    import numpy as np
    from oracle import pipeline
    params = synthetic_params(N // 2 + 1, 0)
    mix, _ = pipeline.synth_mixture(seconds, 1000)
    t = time.perf_counter()
    pipeline.separate(mix, params, "dsd", frameSize=N, hopSize=512, window=np.hanning, overlap=25)
    return time.perf_counter() - t


def main():
    args = parse_args()
    # stdout must carry exactly ONE JSON line: NCCL / torchrun / libraries print banners to fd 1
    # ("NCCL version ..."), so everything else is sent to stderr and the line is written to the
    # original descriptor at the end.
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real, "w")
    global _emit
    _emit = lambda line: (out.write(line + "\n"), out.flush())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
