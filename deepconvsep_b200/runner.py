"""Dataset-level separation runner = the `if not skip_sep:` branch of the reference's trainers
(examples/dsd100/trainCNN.py:285-335, examples/ikala/trainCNN.py:246-285, same shape for Bach10 /
hiphopss): walk the dataset directory, separate every mixture with util's zero-padded patcher and
the analysis window the features were computed with, write the stems with the input's bit depth
(util.writeAudioScipy, util.py:56-58).  The reference does this one file at a time inside the
training process; here the model stays resident on the GPU and the songs are sharded over the
GPUs of a box (one process per GPU, `torchrun`), with no collective on the data path.

    python -m deepconvsep_b200.runner --family dsd --db <DSD100/Mixtures> --out <dir> --model model.pkl
    torchrun --nproc-per-node 8 -m deepconvsep_b200.runner --family dsd --db ... --out ... --model ...
"""
import argparse
import os
import numpy as np

from . import util
from .engine import Separator
from .models import load_model, FAMILY_DEFAULTS
from .sharding import shard_clips, reduce_stats

# trainer settings: (frameSize, hop, window, overlap) -- dsd100/trainCNN.py:431,399; ikala/trainCNN.py:382;
# bach10/trainCNNbach10.py uses 4096 / blackmanharris
TRAINER = {
    "dsd": dict(frameSize=1024, hopSize=512, window="blackmanharris", overlap=25),
    "ikala": dict(frameSize=1024, hopSize=512, window="blackmanharris", overlap=20),
    "bach10": dict(frameSize=4096, hopSize=512, window="blackmanharris", overlap=25),
    # stereo / ILD trainer: transformFFT(frameSize=1024, hopSize=512, window=hanning), overlap 25
    # (dsd100_2ch_ILD/trainCNN_ILD_DSD100.py:487, 438-441)
    "dsd_ild": dict(frameSize=1024, hopSize=512, window="hanning", overlap=25),
}


def list_jobs(family, testdir, outdir):
    """[(input wav, [output wavs])] in the reference's directory conventions."""
    jobs = []
    if family == "dsd_ild":
        # --db is the DSD100 root here: <db>/Mixtures/<sub>/<song>/mixture.wav -> <out>/Sources/<sub>/<song>/<source>.wav
        # (trainCNN_ILD_DSD100.py:296-300, 329-343)
        src = FAMILY_DEFAULTS["dsd_ild"]["sources"]
        for sub in ("Dev", "Test"):
            d = os.path.join(testdir, "Mixtures", sub)
            if not os.path.isdir(d):
                continue
            for f in sorted(os.listdir(d)):
                if f.startswith('.'):
                    continue
                jobs.append((os.path.join(d, f, "mixture.wav"),
                             [os.path.join(outdir, "Sources", sub, f, s + ".wav") for s in src]))
    elif family == "dsd":
        src = FAMILY_DEFAULTS["dsd"]["sources"]
        for sub in ("Dev", "Test"):
            d = os.path.join(testdir, sub)
            if not os.path.isdir(d):
                continue
            for f in sorted(os.listdir(d)):
                if f.startswith('.'):
                    continue
                jobs.append((os.path.join(d, f, "mixture.wav"), [os.path.join(outdir, sub, f, s + ".wav") for s in src]))
    elif family == "ikala":
        for f in sorted(os.listdir(testdir)):
            if f.endswith(".wav"):
                jobs.append((os.path.join(testdir, f), [os.path.join(outdir, f.replace(".wav", "-voice.wav")),
                                                        os.path.join(outdir, f.replace(".wav", "-music.wav"))]))
    else:
        src = FAMILY_DEFAULTS[family]["sources"]
        for f in sorted(os.listdir(testdir)):
            if f.endswith(".wav"):
                jobs.append((os.path.join(testdir, f), [os.path.join(outdir, f.replace(".wav", "_" + s + ".wav")) for s in src]))
    return jobs


def _num_samples(path):
    """Samples per channel from the wav header (memory-mapped, nothing is read); the shard balance and the
    longest-first order go by duration, not by bytes (a 24/32-bit song is not longer than a 16-bit one)."""
    import scipy.io.wavfile
    try:
        _, a = scipy.io.wavfile.read(path, mmap=True)
        return int(a.shape[0])
    except Exception:  # noqa: BLE001  (unreadable header: fall back to the byte count, the read itself will report)
        return int(os.path.getsize(path))


def separate_dataset(family, testdir, outdir, model, scale_factor=0.3, time_context=30, rank=0, world_size=1, device=0,
                     **overrides):
    cfg = dict(TRAINER[family], **overrides)
    params = load_model(model) if isinstance(model, str) else model
    sep = Separator(params, arch=None if family == "ikala" else family, frame_size=cfg["frameSize"], hop=cfg["hopSize"],
                    window=cfg["window"], scale_factor=scale_factor, time_context=time_context, overlap=cfg["overlap"],
                    patcher="util", device=device, feat_size=cfg["frameSize"] // 2 + 1)
    jobs = list_jobs(family, testdir, outdir)
    sizes = [_num_samples(j[0]) for j in jobs]
    seconds = 0.0
    # longest first: the workspace buffers only grow, so the first song sizes them once for the whole shard
    for idx in sorted(shard_clips(sizes, world_size, rank), key=lambda i: (-sizes[i], i)):
        wav, outs = jobs[idx]
        audioObj, sampleRate, bitrate = util.readAudioScipy(wav)
        assert sampleRate == 44100, "Sample rate needs to be 44100"
        if family == "dsd_ild":                              # both channels in, stereo stems out
            assert audioObj.ndim == 2 and audioObj.shape[1] == 2, "the stereo / ILD network needs 2-channel mixtures"
            sep_audio = sep.separate_stereo(audioObj)        # [nsamples, nsrc, 2]
            for i, path in enumerate(outs):
                os.makedirs(os.path.dirname(path), exist_ok=True)
                util.writeAudioScipy(path, sep_audio[:, i, :].astype(np.float64), sampleRate, bitrate)
            seconds += audioObj.shape[0] / float(sampleRate)
            continue
        if audioObj.ndim == 1:
            audio = audioObj
        elif family == "ikala":
            audio = audioObj[:, 0] + audioObj[:, 1]          # ikala/trainCNN.py:255
        else:
            audio = (audioObj[:, 0] + audioObj[:, 1]) / 2    # dsd100/trainCNN.py:304
        stems = sep.separate(audio)
        for i, path in enumerate(outs):
            os.makedirs(os.path.dirname(path), exist_ok=True)
            util.writeAudioScipy(path, stems[i].astype(np.float64), sampleRate, bitrate)
        seconds += len(audio) / float(sampleRate)
    return seconds, len(jobs)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--family", required=True, choices=sorted(TRAINER))
    ap.add_argument("--db", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--model", required=True)
    ap.add_argument("--scale-factor", type=float, default=0.3)
    args = ap.parse_args(argv)
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import time
    t0 = time.time()
    secs, njobs = separate_dataset(args.family, args.db, args.out, args.model, args.scale_factor, rank=rank,
                                   world_size=world, device=local)
    tot, ms, _ = reduce_stats(secs, (time.time() - t0) * 1e3)
    if rank == 0:
        print("separated %d files, %.1f audio-s in %.2f s (%.0f x real time) on %d GPU(s)" % (njobs, tot, ms / 1e3,
                                                                                          tot / (ms / 1e3), world))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
