#!/usr/bin/env python3
"""Stress check for the wavefront / multi-stream synchronisation: the 600-frame 1080p clip must encode to the
same bytes on every run and for every number of GOPs in flight, and libavcodec must decode it."""
import hashlib, os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cedarx_h264_encoder_b200 as cx
from cedarx_h264_encoder_b200 import api, synth
w, h, n, gop = 1920, 1088, int(os.environ.get("N", 600)), 60
cabac = int(os.environ.get("CABAC", 1))
ref = None
for lanes, reps in ((0, 3), (1, 1), (3, 1)):
    enc = cx.Encoder(api.make_config(w, h, qp=25, gop=gop, cabac=cabac, max_clip_frames=n, gops_in_flight=lanes))
    st = torch.from_numpy(enc.clip_input(n))
    for i in range(0, n, 20):
        st[i:i + 20].copy_(synth.synth_clip(w, h, list(range(i, min(n, i + 20))), 0, device="cuda"))
    torch.cuda.synchronize()
    for r in range(reps):
        enc.clip_upload(n); enc.clip_encode(n, 0)
        data, sizes = enc.clip_download(n)
        hsh = hashlib.sha256(data.tobytes()).hexdigest()
        print("lanes", lanes, "rep", r, "bytes", len(data), hsh[:16])
        if ref is None:
            ref = hsh
            if os.environ.get("DECODE", "1") == "1":
                import avdec
                dec = avdec.decode(data[:int(sizes[:3].sum())].tobytes())
                print("decoded", len(dec), "frames of the first 3")
        assert hsh == ref, "NON-DETERMINISTIC OUTPUT"
    enc.close()
print("STRESS OK")
