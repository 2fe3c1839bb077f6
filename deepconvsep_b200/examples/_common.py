"""Shared driver of the stand-alone separation scripts: `train_auto` of
examples/dsd100/separate_dsd.py:239-313 (and its iKala / Bach10 siblings) on the CUDA pipeline."""
import os
import threading
import numpy as np
import scipy.io.wavfile

from ..engine import Separator
from ..models import load_model, FAMILY_DEFAULTS

_cache = {}
_cache_lock = threading.Lock()       # run_many's worker threads share the cache


def get_separator(model, arch, frame_size, hop, window, scale_factor, time_context, overlap, feat_size, device=0, slot=0):
    key = (os.path.abspath(model), os.path.getmtime(model), arch, frame_size, hop, str(window), scale_factor,
           time_context, overlap)
    with _cache_lock:
        if _cache.get("key") != key:
            _cache.clear()   # one resident model at a time (Bach10 weights are 856 MB), per (device, slot)
            _cache["key"] = key
        if (device, slot) not in _cache:
            _cache[(device, slot)] = Separator(load_model(model), arch=arch, frame_size=frame_size, hop=hop, window=window,
                                               scale_factor=scale_factor, time_context=time_context, overlap=overlap,
                                               patcher="standalone", feat_size=feat_size, device=device)
        return _cache[(device, slot)]


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
        out_name, window=None, device=0, slot=0):
    """wav in -> one int16 wav per source in `outdir`.  `batch_size` is accepted for signature
    compatibility; the CUDA path has no patch batches."""
    d = dict(FAMILY_DEFAULTS[family])
    if window is not None:
        d["window"] = window
    sampleRate, audioObj = scipy.io.wavfile.read(filein)
    if sampleRate != 44100:
        print("Sample rate is not 44100")        # separate_dsd.py:313
        return None
    arch = None if family in ("ikala",) else family
    if isinstance(device, (list, tuple)):
        # one recording over several GPUs: hop- and patch-aligned segments with margins, one host thread per device,
        # the stitched stems are those of the whole-clip call (deepconvsep_b200.longclip)
        from .. import longclip
        seps = [get_separator(model, arch, frame_size, hop, d["window"], scale_factor, time_context, overlap, input_size,
                              device=dev, slot=slot) for dev in device]
        sep = seps[0]
        stems = longclip.separate_long(seps, decode(audioObj, family))
        stems16 = (stems.astype(np.float64) * np.iinfo(np.int16).max).astype('int16')
    elif audioObj.dtype == np.int16 and family != "ikala":
        sep = get_separator(model, arch, frame_size, hop, d["window"], scale_factor, time_context, overlap, input_size,
                            device=device, slot=slot)
        stems16 = sep.separate_pcm16(audioObj, downmix=1)          # decode/downmix/encode on the GPU
    else:
        sep = get_separator(model, arch, frame_size, hop, d["window"], scale_factor, time_context, overlap, input_size,
                            device=device, slot=slot)
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


# ---- command line shared by the separate_*.py scripts ------------------------------------------------------
LONG_OPTS = ["ifile=", "odir=", "mfile=", "frame-size=", "window=", "devices=", "batch-clips="]
EXTRA_USAGE = ("  optional: --frame-size N (STFT frame, feat_size = N/2+1)  --window hanning|blackmanharris|sinebell\n"
               "            --devices 0,1,...  --batch-clips K (clips in flight per device); with these, -i may be a directory of wavs\n"
               "            (one wav and several devices: the recording itself is cut into segments over the devices)")


def parse_cli(argv, usage):
    """getopt like the reference scripts (`-i -o -m`, separate_dsd.py:316-332) plus the long options SURVEY.md 5 asks for."""
    import getopt
    import sys
    try:
        opts, _ = getopt.getopt(argv, "hi:o:m:", LONG_OPTS)
    except getopt.GetoptError:
        print(usage)
        print(EXTRA_USAGE)
        sys.exit(2)
    o = {"inputfile": None, "outdir": None, "model": None, "frame_size": None, "window": None, "devices": None, "batch_clips": 1}
    for opt, arg in opts:
        if opt == "-h":
            print(usage)
            print(EXTRA_USAGE)
            sys.exit()
        elif opt in ("-i", "--ifile"):
            o["inputfile"] = arg
        elif opt in ("-o", "--odir"):
            o["outdir"] = arg
        elif opt in ("-m", "--mfile"):
            o["model"] = arg
        elif opt == "--frame-size":
            o["frame_size"] = int(arg)
        elif opt == "--window":
            o["window"] = arg
        elif opt == "--devices":
            o["devices"] = [int(x) for x in arg.split(",") if x != ""]
        elif opt == "--batch-clips":
            o["batch_clips"] = max(1, int(arg))
    if o["inputfile"] is None or o["outdir"] is None or o["model"] is None:
        print(usage)
        sys.exit(2)
    return o


def cli_main(argv, usage, train_auto_default, run_one):
    """`train_auto_default(inputfile, outdir, model)` = the script's literal reference call (no extra flag given);
    `run_one(filein, outdir, model, frame_size, window, device, slot, several_clips)` = the same with the overrides."""
    o = parse_cli(argv, usage)
    plain = o["frame_size"] is None and o["window"] is None and o["devices"] is None and o["batch_clips"] == 1 \
        and not os.path.isdir(o["inputfile"])
    if plain:
        return train_auto_default(o["inputfile"], o["outdir"], o["model"])
    if os.path.isdir(o["inputfile"]):
        files = sorted(os.path.join(o["inputfile"], f) for f in os.listdir(o["inputfile"]) if f.lower().endswith(".wav"))
    else:
        files = [o["inputfile"]]
    devices = o["devices"] or [0]
    if len(files) == 1 and len(devices) > 1:
        # a single recording and several GPUs: split the recording (run() with a device list), not the file list
        return [run_one(files[0], o["outdir"], o["model"], o["frame_size"], o["window"], devices, 0, False)]
    return run_many(files, o["outdir"], o["model"], o["frame_size"], o["window"], devices, o["batch_clips"], run_one)


def run_many(files, outdir, model, frame_size, window, devices, batch_clips, run_one):
    """The reference's only multi-clip driver spawns one Python process per file (separate_multiple.ipynb cell 3); here the
    clips go, longest first, to `len(devices) x batch_clips` resident pipelines (one context / stream each)."""
    import queue
    import threading
    order = sorted(files, key=lambda f: -os.path.getsize(f))
    q = queue.Queue()
    for f in order:
        q.put(f)
    results, errors = {}, []

    def worker(device, slot):
        import torch
        torch.cuda.set_device(device)
        with torch.cuda.stream(torch.cuda.Stream(device=device)):
            while True:
                try:
                    f = q.get_nowait()
                except queue.Empty:
                    return
                try:
                    results[f] = run_one(f, outdir, model, frame_size, window, device, slot, len(files) > 1)
                except Exception as e:  # noqa: BLE001  (reported after the pool drains)
                    errors.append((f, e))
    ths = [threading.Thread(target=worker, args=(d, s)) for d in devices for s in range(batch_clips)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errors:
        raise errors[0][1]
    return [results[f] for f in files]
