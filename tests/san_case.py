#!/usr/bin/env python3
"""Small cases that run every kernel (both entropy coders, I + P, slices, Intra4x4, intra in P, NV16, ragged size, clip and
frame mode, the pipeline) for `compute-sanitizer --tool memcheck python tests/san_case.py`; checks the bytes against the
golden model as it goes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cedarx_h264_encoder_b200 as cx
from cedarx_h264_encoder_b200 import api
from common import make_clip, oracle_encode_clip, split_frame

cases = [dict(w=96, h=80, fmt=0, cabac=1), dict(w=96, h=80, fmt=0, cabac=0), dict(w=86, h=50, fmt=0, cabac=1, slice_rows=2),
         dict(w=64, h=48, fmt=1, cabac=1), dict(w=96, h=80, fmt=0, cabac=1, intra4x4=1, p_intra=1), dict(w=208, h=160, fmt=0, cabac=1, me=64)]
for c in cases:
    w, h, fmt, n, gop = c["w"], c["h"], c["fmt"], 7, 3
    kw = dict(qp=24, gop=gop, cabac=c["cabac"], me_range=c.get("me", 8), slice_rows=c.get("slice_rows", 0),
              intra4x4=c.get("intra4x4", 0), p_intra=c.get("p_intra", 0))
    clip = make_clip("synth", w, h, n, fmt)
    want, sizes, _ = oracle_encode_clip(clip, w, h, fmt, **kw)
    with cx.Encoder(api.make_config(w, h, fmt=fmt, max_clip_frames=n, gops_in_flight=2, **kw)) as enc:
        got, _ = enc.encode_clip(clip)
    assert got == want, c
    with cx.Encoder(api.make_config(w, h, fmt=fmt, **kw)) as enc:
        frames = b"".join(enc.encode(*split_frame(clip[t], w, h, fmt)) for t in range(n))
    assert frames == want, c
    print("ok", c, flush=True)
clip = make_clip("synth", 96, 80, 14)
want, _, _ = oracle_encode_clip(clip, 96, 80, qp=24, gop=3, cabac=1, me_range=8)
with cx.Pipe(api.make_config(96, 80, qp=24, gop=3, cabac=1, me_range=8), None, 2, 1) as pipe:
    got, _ = pipe.encode(clip)
assert got == want
print("SANITIZER CASES OK")
