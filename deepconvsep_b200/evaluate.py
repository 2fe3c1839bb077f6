"""Dataset-level evaluation = the reference's MATLAB drivers with the BSS-Eval work on the GPU
(deepconvsep_b200/evaluation.py):

  iKala    evaluation/evaluate_SS_iKala.m:1-70   per file: SDR / SIR / SAR of [voice, music] with
           bss_eval_sources on gain-normalised signals, and the normalised NSDR / NSIR / NSAR = the
           improvement over using the mixture as both estimates; saved as <file>.mat with the
           reference's variable names
  Bach10   evaluation/Bach10_eval_only.m:94      bss_eval_sources on the four instruments
  DSD100   evaluation/DSD100_eval_only.m:110-215 per song: windowed (30 s / 15 s) multichannel
           bss_eval_images of the four sources and of [vocals, accompaniment]; saved as
           <song>_results.mat with a `results` struct (name, <source>.sdr/.isr/.sir/.sar)

    python -m deepconvsep_b200.evaluate ikala  --root <iKala dir> --method <name>
    python -m deepconvsep_b200.evaluate dsd100 --dataset <DSD100 dir> --estimates <dir>/<method>
"""
import argparse
import os
import numpy as np
import scipy.io

from . import evaluation, util

DEVICE = "cuda"      # tests point this at "cpu" together with lag-free stand-ins for the metric


def _dev(x):
    import torch
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32), device=DEVICE)


def _read(path):
    audio, fs, _ = util.readAudioScipy(path)       # float in [-1, 1], [nsampl] or [nsampl, nchan]
    return np.asarray(audio, dtype=np.float64), fs


# ------------------------------------------------------------------------------------------ iKala
def evaluate_ikala_file(source_wav, est_voice_wav, est_music_wav, out_mat=None, flen=evaluation.FLEN):
    """evaluate_SS_iKala.m:40-66 for one file -> dict with SDR, SIR, SAR, NSDR, NSIR, NSAR ([voice, music])"""
    src, _ = _read(source_wav)
    voice, karaoke = src[:, 1], src[:, 0]          # :47-48: channel 2 is the voice, channel 1 the music
    mixed = (voice + karaoke) / 2
    ev, _ = _read(est_voice_wav)
    ek, _ = _read(est_music_wav)
    ev, ek = (x if x.ndim == 1 else x[:, 0] for x in (ev, ek))
    ev, ek = ev[:len(voice)], ek[:len(karaoke)]     # :52-57
    n = min(len(ev), len(voice))
    ref = np.stack([voice[:n], karaoke[:n]]) / np.linalg.norm(voice[:n] + karaoke[:n])
    est = np.stack([ev[:n], ek[:n]]) / np.linalg.norm(ev[:n] + ek[:n])
    mix = np.stack([mixed[:n], mixed[:n]]) / np.linalg.norm(2 * mixed[:n])
    sdr, sir, sar, _ = evaluation.bss_eval_sources(_dev(est), _dev(ref), flen=flen)
    nsdr, nsir, nsar, _ = evaluation.bss_eval_sources(_dev(mix), _dev(ref), flen=flen)
    res = {"SDR": sdr, "SIR": sir, "SAR": sar, "NSDR": sdr - nsdr, "NSIR": sir - nsir, "NSAR": sar - nsar}
    if out_mat:
        os.makedirs(os.path.dirname(out_mat) or ".", exist_ok=True)
        scipy.io.savemat(out_mat, {k: v.reshape(-1, 1) for k, v in res.items()})
    return res


def evaluate_ikala(root, method, flen=evaluation.FLEN):
    """every <root>/Wavfile/*.wav with estimates <root>/output/<method>/<file>-voice.wav / -music.wav ->
    <root>/measures/test_<method>/<file>.mat (evaluate_SS_iKala.m:22-39); existing results are kept"""
    done = []
    wavdir, estdir = os.path.join(root, "Wavfile"), os.path.join(root, "output", method)
    outdir = os.path.join(root, "measures", "test_" + method)
    for f in sorted(os.listdir(wavdir)):
        if not f.endswith(".wav"):
            continue
        ev, ek = (os.path.join(estdir, f.replace(".wav", "-%s.wav" % s)) for s in ("voice", "music"))
        mat = os.path.join(outdir, f.replace(".wav", ".mat"))
        if os.path.isfile(ev) and not os.path.isfile(mat):
            evaluate_ikala_file(os.path.join(wavdir, f), ev, ek, mat, flen)
            done.append(mat)
    return done


# ------------------------------------------------------------------------------------------ Bach10
def evaluate_sources(est_wavs, ref_wavs, flen=evaluation.FLEN):
    """bss_eval_sources on lists of mono wav files (Bach10_eval_only.m:94) -> (SDR, SIR, SAR, perm)"""
    est = [(_read(p)[0]) for p in est_wavs]
    ref = [(_read(p)[0]) for p in ref_wavs]
    mono = lambda x: x if x.ndim == 1 else x.mean(axis=1)
    n = min(min(len(x) for x in est), min(len(x) for x in ref))
    return evaluation.bss_eval_sources(_dev(np.stack([mono(x)[:n] for x in est])),
                                       _dev(np.stack([mono(x)[:n] for x in ref])), flen=flen)


# ------------------------------------------------------------------------------------------ DSD100
DSD_SOURCES = ["bass", "drums", "other", "vocals"]                         # DSD100_eval_only.m:82
DSD_ESTIMATE_FILES = ["mixture_bass", "mixture_drums", "mixture_others", "mixture_vocals", "mixture_accompaniment"]


def _stereo(x):
    return np.repeat(x[:, None], 2, axis=1) if x.ndim == 1 else (np.repeat(x, 2, axis=1) if x.shape[1] == 1 else x)


def evaluate_dsd100_song(sources_song_dir, estimates_song_dir, out_mat=None, estimate_files=None, win_s=30, hop_s=15,
                         flen=evaluation.FLEN):
    """DSD100_eval_only.m:118-196 for one song: windowed multichannel BSS-Eval of bass / drums / other /
    vocals and of the accompaniment (sum of the first three; an estimate file if present, else the sum of
    the three estimates) -> {source: {sdr, isr, sir, sar}} with one value per window"""
    files = estimate_files or DSD_ESTIMATE_FILES
    refs, fs = [], None
    for s in DSD_SOURCES:
        x, fs = _read(os.path.join(sources_song_dir, s + ".wav"))
        refs.append(_stereo(x))                                             # mono -> duplicated (:130)
    nsampl = refs[0].shape[0]
    ests = [np.zeros((nsampl, 2)) for _ in range(4)]
    for k in range(4):
        p = os.path.join(estimates_song_dir, files[k] + ".wav")
        if os.path.isfile(p):
            e = _stereo(_read(p)[0])
            nsampl = min(nsampl, e.shape[0])                                 # :149-151
            ests[k] = e
    refs = [r[:nsampl] for r in refs]
    ests = [e[:nsampl] if e.shape[0] >= nsampl else np.pad(e, ((0, nsampl - e.shape[0]), (0, 0))) for e in ests]
    acc_ref = refs[0] + refs[1] + refs[2]
    pacc = os.path.join(estimates_song_dir, files[4] + ".wav")
    if os.path.isfile(pacc):
        acc_est = _stereo(_read(pacc)[0])[:nsampl]
        nsampl = min(nsampl, acc_est.shape[0])
    else:
        acc_est = ests[0] + ests[1] + ests[2]                                # :166
    cut = lambda xs: np.stack([x[:nsampl].T for x in xs])                    # -> [nsrc, nchan, nsampl]
    win, ove = int(win_s * fs), int(hop_s * fs)
    src = evaluation.bss_eval_windowed(_dev(cut(ests)), _dev(cut(refs)), win, ove, flen=flen)
    acc = evaluation.bss_eval_windowed(_dev(cut([ests[3], acc_est])), _dev(cut([refs[3], acc_ref])), win, ove, flen=flen)
    results = {"name": os.path.basename(os.path.normpath(sources_song_dir))}
    for k, s in enumerate(DSD_SOURCES):
        results[s] = {m: src[q][k] for q, m in enumerate(("sdr", "isr", "sir", "sar"))}
    results["accompaniment"] = {m: acc[q][1] for q, m in enumerate(("sdr", "isr", "sir", "sar"))}   # :180-183
    if out_mat:
        os.makedirs(os.path.dirname(out_mat) or ".", exist_ok=True)
        scipy.io.savemat(out_mat, {"results": results})
    return results


def evaluate_dsd100(dataset, estimates, subsets=("Test", "Dev"), **kw):
    """every song of <dataset>/Sources/<subset> with estimates in <estimates>/<subset>/<song> ->
    <estimates>/<subset>/<song>_results.mat (DSD100_eval_only.m:100-116, 198-203); existing results are kept"""
    done = []
    for sub in subsets:
        sdir, edir = os.path.join(dataset, "Sources", sub), os.path.join(estimates, sub)
        if not os.path.isdir(sdir):
            continue
        for song in sorted(os.listdir(sdir)):
            mat = os.path.join(edir, song + "_results.mat")
            if song.startswith(".") or os.path.isfile(mat) or not os.path.isdir(os.path.join(edir, song)):
                continue
            evaluate_dsd100_song(os.path.join(sdir, song), os.path.join(edir, song), mat, **kw)
            done.append(mat)
    return done


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("ikala")
    a.add_argument("--root", required=True)
    a.add_argument("--method", required=True)
    b = sub.add_parser("dsd100")
    b.add_argument("--dataset", required=True)
    b.add_argument("--estimates", required=True)
    args = ap.parse_args(argv)
    done = evaluate_ikala(args.root, args.method) if args.cmd == "ikala" else evaluate_dsd100(args.dataset, args.estimates)
    print("wrote %d result files" % len(done))


if __name__ == "__main__":
    main()
