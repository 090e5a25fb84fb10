#!/usr/bin/env python3
"""Smallest program that runs every kernel of the clip path at the benchmark's geometry (1920x1088, ten GOPs in lock
step): `frames_per_gop` steps of 10 lanes.  For ncu captures (-k regex:<kernel>).
    python tools/prof_clip.py [frames_per_gop=4] [repeats=1] [lanes=10]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cedarx_h264_encoder_b200 as cx  # noqa: E402
from cedarx_h264_encoder_b200 import api, synth  # noqa: E402

gop = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 10
w, h, n = 1920, 1088, gop * lanes
ids = [g * 60 + t for g in range(lanes) for t in range(gop)]  # the first frames of GOPs of the benchmark clip
enc = cx.Encoder(api.make_config(w, h, qp=25, gop=gop, cabac=1, max_clip_frames=n, gops_in_flight=lanes))
enc.clip_input(n)[:] = synth.synth_clip(w, h, ids, 0, device="cuda").cpu().numpy()
enc.clip_upload(n)
for _ in range(reps):
    enc.clip_encode(n, 0)
data, sizes = enc.clip_download(n)
torch.cuda.synchronize()
print("encoded %d frames, %d bytes" % (n, len(data)))
enc.close()
