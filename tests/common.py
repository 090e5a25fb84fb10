"""Shared helpers of the test-suite: deterministic contents and oracle-side clip encoding."""
import numpy as np

import oracle_lib as O


def content(kind, w, h, t, fmt=0):
    crows = h if fmt else h // 2
    if kind == "synth":
        return O.synth_frame(w, h, t, fmt)
    if kind == "noise":
        r = np.random.default_rng(100 + t)
        return r.integers(0, 256, (h, w), dtype=np.uint8), r.integers(0, 256, (crows, w), dtype=np.uint8)
    if kind == "static":  # mostly P_Skip
        y, c = O.synth_frame(w, h, 0, fmt)
        if t % 3 == 2:
            y = y.copy()
            y[16:48, 32:80] = O.synth_frame(w, h, t, fmt)[0][16:48, 32:80]
        return y, c
    if kind == "shift":  # translating noise field: large exact motion vectors, clamped borders
        r = np.random.default_rng(7)
        big = r.integers(0, 256, (h + 400, w + 400), dtype=np.uint8)
        bigc = r.integers(0, 256, (crows + 200, w + 400), dtype=np.uint8)
        ox, oy = 100 + 7 * t, 100 - 3 * t
        return big[oy:oy + h, ox:ox + w].copy(), bigc[oy // 2:oy // 2 + crows, (ox // 2) * 2:(ox // 2) * 2 + w].copy()
    if kind == "flat":
        return np.full((h, w), (37 * t) % 256, np.uint8), np.full((crows, w), 128, np.uint8)
    raise ValueError(kind)


def make_clip(kind, w, h, n, fmt=0):
    fb = w * h * (2 if fmt else 3) // (1 if fmt else 2)
    clip = np.empty((n, fb), np.uint8)
    for t in range(n):
        y, c = content(kind, w, h, t, fmt)
        clip[t, :w * h] = y.reshape(-1)
        clip[t, w * h:] = c.reshape(-1)
    return clip


def split_frame(clip_row, w, h, fmt=0):
    crows = h if fmt else h // 2
    return clip_row[:w * h].reshape(h, w), clip_row[w * h:].reshape(crows, w)


def oracle_encode_clip(clip, w, h, fmt=0, keep_recon=False, **cfg):
    enc = O.Encoder(O.make_config(w, h, fmt=fmt, **cfg))
    stream, sizes, recs = b"", [], []
    for row in clip:
        y, c = split_frame(row, w, h, fmt)
        b = enc.encode(y, c)
        stream += b
        sizes.append(len(b))
        if keep_recon:
            recs.append(enc.recon())
    enc.close()
    return stream, sizes, recs
