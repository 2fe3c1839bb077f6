"""One long recording over several GPUs ("for a single very long clip, split by frame
range with halo").  The reference has no such path -- `train_auto` holds the whole clip's batches in
host memory (examples/dsd100/separate_dsd.py:289-300) -- but every stage of the path is local in time,
so a recording can be cut into sample ranges that are separated independently and stitched, and the
result is the one the whole-clip pipeline produces:

* STFT frame t reads samples [tH - N/2, tH + N/2) (transform.py:309-332);
* the network sees patches of `time_context` frames on the grid k * step, step = time_context - overlap
  (separate_dsd.py:114-135, util.py:220-248);
* the sequential cross-fade makes spectrum frame f a function of the patches that cover it, from the one
  that overwrites it (offset >= overlap) onwards (separate_dsd.py:155-168);
* the inverse STFT overlap-adds the N/H frames around a sample, and its normaliser is the sum of the
  window products of the same frames (transform.py:379-394).

A segment therefore starts on a hop boundary that is also a patch boundary of the whole clip's grid
(frame0 a multiple of step), carries a left margin long enough that the first sample it contributes has
only frames of patches free of the sub-clip's own front zero padding, and a right margin for the
mirror-image condition (plan_segments documents the two bounds).  The first segment starts at sample 0 and
the last ends at the clip's end, so the true edges are reproduced by the pipeline itself.  What a segment
computes in its margins is discarded.

The arithmetic of a kept sample is that of the whole-clip run up to the summation order inside the GEMMs
(the split of K over CTAs depends on the number of patches), i.e. to float32 rounding; tests/ check the
stitched result against the whole-clip oracle at the north-star tolerance and the planner against the
oracle exactly (float64: bit-identical stitching)."""
import threading
from collections import namedtuple
import numpy as np

Segment = namedtuple("Segment", "in_start in_stop out_start out_stop frame0")


def _ceil_div(a, b):
    return -((-a) // b)


def margins(frame_size, hop, time_context, overlap):
    """(left, right) margins in samples that plan_segments asks for (before rounding to the grids)."""
    step = time_context - overlap
    q = _ceil_div(frame_size // 2, hop)                    # frames touched by the sub-clip's zero padding at either end
    s_v = _ceil_div(q, step) * step                        # first patch of the sub-clip without such a frame
    left = (s_v + overlap) * hop + frame_size // 2
    right = (q + time_context - 1) * hop + frame_size // 2 + hop
    return left, right


def plan_segments(num_samples, parts, frame_size, hop, time_context, overlap):
    """Cut [0, num_samples) into at most `parts` segments.  Each Segment holds the sample range to feed the pipeline
    (in_start:in_stop), the range of the result that is kept (out_start:out_stop, absolute sample indices) and the
    index of the whole-clip STFT frame its first frame corresponds to (a multiple of step; score filters are sliced
    from there).

    Left bound.  With s0 = frame0 * hop the sub-clip's frame t' is the clip's frame frame0 + t' once
    t' >= q = ceil(N/2 / hop) (no front padding inside the frame); the first patch made of such frames starts at
    s_v = ceil_step(q); the cross-fade gives the clip's value to spectrum frames f' >= s_v + overlap; a sample n' of
    the sub-clip is summed from frames t' > (n' - N/2) / hop, so n' >= (s_v + overlap) * hop + N/2 is exact.
    Right bound.  With G = (in_stop - in_start) / hop, frames t' <= G - q are free of the back padding; the last patch
    made of such frames starts at s_l = floor_step(G - q - time_context + 1); every patch covering f' < s_l + step
    exists in both runs and is exact; sample n' is summed from frames t' <= (n' + N/2) / hop, so
    n' < (s_l + step) * hop - N/2 is exact; G >= ceil((n'_stop + N/2) / hop) + q + time_context - 1 guarantees it."""
    L, N, H = int(num_samples), int(frame_size), int(hop)
    step = time_context - overlap
    assert step > 0 and L >= 0 and parts >= 1
    q = _ceil_div(N // 2, H)
    s_v = _ceil_div(q, step) * step
    left, right = margins(N, H, time_context, overlap)
    # no point in cores shorter than the margins they drag along
    parts = max(1, min(int(parts), L // max(1, 2 * (left + right))))
    cuts = [_ceil_div(L * r, parts * H) * H for r in range(parts)] + [L]       # cores start on hop boundaries
    segs = []
    for r in range(parts):
        o0, o1 = cuts[r], cuts[r + 1]
        if o1 <= o0:
            continue
        # largest step-aligned frame0 with o0 - frame0*H >= (s_v + overlap)*H + N/2
        g0 = ((o0 - N // 2) // H - s_v - overlap) // step * step
        if r == 0 or g0 <= 0:      # the margin reaches the clip's start: the true edge is reproduced by the pipeline itself
            g0 = 0
        s0 = g0 * H
        if r == parts - 1:
            s1 = L
        else:
            G = _ceil_div(o1 - s0 + N // 2, H) + q + time_context - 1
            s1 = s0 + G * H
            if s1 >= L:
                s1 = L
        segs.append(Segment(s0, s1, o0, o1, g0))
    return segs


def stitch(segments, pieces, num_samples, dtype=np.float32):
    """pieces[i]: array [..., in_stop - in_start] of segment i (sample axis last) -> [..., num_samples]."""
    lead = pieces[0].shape[:-1]
    out = np.zeros(lead + (int(num_samples),), dtype=dtype)
    for sg, p in zip(segments, pieces):
        assert p.shape[-1] == sg.in_stop - sg.in_start, (p.shape, sg)
        out[..., sg.out_start:sg.out_stop] = p[..., sg.out_start - sg.in_start:sg.out_stop - sg.in_start]
    return out


def _geometry(sep):
    return sep.frame_size, sep.hop, sep.model.tc, sep.overlap


def _run(sep, sub, filt):
    """One segment through a Separator (or a callable (sub, filt) -> array with the sample axis last)."""
    if not hasattr(sep, "model"):
        return np.asarray(sep(sub, filt))
    arch = sep.model.arch
    if arch == "bach10_score":
        return sep.separate_score(sub, filt)
    if arch == "dsd_ild":
        return np.ascontiguousarray(sep.separate_stereo(sub).transpose(1, 2, 0))     # [L, nsrc, 2] -> [nsrc, 2, L]
    return sep.separate(sub)


def _slice_filters(filters, sg, hop):
    if filters is None:
        return None
    T = _ceil_div(sg.in_stop - sg.in_start, hop) + 2            # transform.py:309
    f = filters[:, sg.frame0:sg.frame0 + T]
    assert f.shape[1] == T, (f.shape, T, sg)
    return f


def separate_long(separators, audio, parts=None, filters=None, geometry=None):
    """audio float [L] (stereo / ILD network: [L, 2]) -> what the Separator's own call returns for the whole clip
    (float32 [nsrc, L]; stereo: [L, nsrc, 2]).  `separators`: one Separator or a list (one per GPU, or several
    contexts of one GPU); segment i runs on separators[i % len], one host thread per separator (the C-ABI calls
    release the GIL).  parts defaults to len(separators).  filters: score filters [4, T, F] of the whole clip
    (score-informed network).  geometry=(frame_size, hop, time_context, overlap) is needed only when the
    separators are plain callables (tests)."""
    seps = list(separators) if isinstance(separators, (list, tuple)) else [separators]
    N, H, tc, ov = geometry if geometry is not None else _geometry(seps[0])
    a = np.asarray(audio)
    L = a.shape[0]
    segs = plan_segments(L, parts or len(seps), N, H, tc, ov)
    pieces = [None] * len(segs)
    errors = []

    def work(w):
        try:
            for i in range(w, len(segs), len(seps)):
                sg = segs[i]
                pieces[i] = _run(seps[w], a[sg.in_start:sg.in_stop], _slice_filters(filters, sg, H))
        except BaseException as e:          # surfaced in the caller's thread
            errors.append(e)

    nthreads = min(len(seps), len(segs))
    if nthreads <= 1:
        work(0)
    else:
        ts = [threading.Thread(target=work, args=(w,)) for w in range(nthreads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    if errors:
        raise errors[0]
    out = stitch(segs, pieces, L, dtype=pieces[0].dtype)
    if hasattr(seps[0], "model") and seps[0].model.arch == "dsd_ild":
        out = np.ascontiguousarray(out.transpose(2, 0, 1))
    return out


def separate_long_distributed(separator, audio, filters=None, geometry=None, group=None):
    """The same over the ranks of an initialised process group (one process per GPU, every rank holds the clip):
    rank r separates segments r, r + world, ...; the kept parts are gathered to rank 0 (sharding.gather_stems,
    the path's only exchange, off the data path) which returns the stitched stems; other ranks return None."""
    import torch.distributed as dist
    from .sharding import gather_stems
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    N, H, tc, ov = geometry if geometry is not None else _geometry(separator)
    a = np.asarray(audio)
    L = a.shape[0]
    segs = plan_segments(L, world, N, H, tc, ov)
    mine = []
    for i in range(rank, len(segs), world):
        sg = segs[i]
        p = _run(separator, a[sg.in_start:sg.in_stop], _slice_filters(filters, sg, H))
        mine.append((i, np.ascontiguousarray(p[..., sg.out_start - sg.in_start:sg.out_stop - sg.in_start])))
    gathered = gather_stems(mine, world, rank, group=group)
    if rank != 0:
        return None
    kept = dict(kv for part in gathered for kv in part)
    first = kept[0]
    out = np.zeros(first.shape[:-1] + (L,), dtype=first.dtype)
    for i, sg in enumerate(segs):
        out[..., sg.out_start:sg.out_stop] = kept[i]
    if hasattr(separator, "model") and separator.model.arch == "dsd_ild":
        out = np.ascontiguousarray(out.transpose(2, 0, 1))
    return out
