#!/usr/bin/env python3
"""Diagnosis: encode one clip with per-launch events and dump the timeline (CEDAR_B200_TIMELINE)."""
import ctypes as C
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = os.path.join(ROOT, "gpurun_out", "timeline.csv")
os.makedirs(os.path.dirname(out), exist_ok=True)
if os.path.exists(out):
    os.remove(out)
os.environ["CEDAR_B200_TIMELINE"] = out
import cedarx_h264_encoder_b200 as cx  # noqa: E402
from cedarx_h264_encoder_b200 import api, synth  # noqa: E402

w, h, n, gop = 1920, 1088, 600, 60
enc = cx.Encoder(api.make_config(w, h, qp=25, gop=gop, cabac=1, max_clip_frames=n))
st = torch.from_numpy(enc.clip_input(n))
for i in range(0, n, 20):
    st[i:i + 20].copy_(synth.synth_clip(w, h, list(range(i, min(n, i + 20))), 0, device="cuda"))
torch.cuda.synchronize()
enc.clip_upload(n)
enc.clip_encode(n, 0)
enc.clip_download(n)
bins = np.zeros(n, np.uint32)
enc.L.cedar_b200_debug_read(enc.h, 6, bins.ctypes.data, bins.nbytes)
print("bins per frame: I", bins[::gop].mean(), "P", np.delete(bins, np.arange(0, n, gop)).mean(), "total", bins.sum())
enc.profile_enable(True)
enc.clip_encode(n, 0)
enc.profile_read()
enc.profile_enable(False)
enc.close()
rows = [l.strip().split(",") for l in open(out)]
import collections
agg = collections.OrderedDict()
for name, a, b in rows:
    a, b = float(a), float(b)
    d = agg.setdefault(name, [1e9, 0, 0, 0.0])
    d[0] = min(d[0], a); d[1] = max(d[1], b); d[2] += 1; d[3] += b - a
for k, v in agg.items():
    print("%-24s first start %8.2f  last end %8.2f  n %4d  sum %9.2f" % (k, v[0], v[1], v[2], v[3]))
side = [(float(a), float(b)) for nme, a, b in rows if nme in ("cabac_kernel", "cabac_resolve_kernel", "cabac_code_kernel")]
print("cabac (start,end) first 6:", [(round(a, 1), round(b, 1)) for a, b in side[:6]], "last 4:", [(round(a, 1), round(b, 1)) for a, b in side[-4:]])
me = [(float(a), float(b)) for nme, a, b in rows if nme == "me_kernel"]
print("me first 4:", [(round(a, 1), round(b, 1)) for a, b in me[:4]], "last 2:", [(round(a, 1), round(b, 1)) for a, b in me[-2:]])
