#!/usr/bin/env python3
"""Experiment: where does the time of a chain go when one 600-frame clip is split over k handles (lane groups)?
Per-kernel live times (CUDA events on the launching streams) per handle, for k = 1, 2, 5 handles."""
import os, sys, time, threading
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cedarx_h264_encoder_b200 as cx
from cedarx_h264_encoder_b200 import api, synth

w, h, gop = 1920, 1080, 60


def make(n, lanes):
    enc = cx.Encoder(api.make_config(w, h, qp=25, gop=gop, cabac=int(os.environ.get("CABAC", 1)), max_clip_frames=n, gops_in_flight=lanes))
    st = torch.from_numpy(enc.clip_input(n))
    for i in range(0, n, 20):
        st[i:i + 20].copy_(synth.synth_clip(w, h, list(range(i, min(n, i + 20))), 0, device="cuda"))
    torch.cuda.synchronize()
    enc.clip_upload(n)
    return enc


for k, n, lanes in [(1, 600, 10), (1, 300, 5), (2, 300, 5), (5, 120, 2)]:
    encs = [make(n, lanes) for _ in range(k)]
    for e in encs:
        e.clip_encode(n, 0)
    torch.cuda.synchronize()
    for mode in (0, 1):
        for e in encs:
            e.profile_enable(mode)
            if mode:
                e.profile_read(reset=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=lambda e=e: [e.clip_encode(n, 0) for _ in range(3)]) for e in encs]
        [t.start() for t in th]; [t.join() for t in th]
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        if mode == 0:
            print("== %d handles x %d frames (%d lanes): %.1f ms per pass of all handles = %.0f frames/s" % (k, n, lanes, dt * 1e3, k * n / dt))
        else:
            p = encs[0].profile_read(reset=True)
            print("   with per-launch events: %.1f ms; handle 0 live ms per clip:" % (dt * 1e3),
                  ", ".join("%s %.1f" % (nm.replace("_kernel", ""), v[0] / 3) for nm, v in p.items() if v[0] / 3 > 0.5))
    for e in encs:
        e.profile_enable(0)
        e.close()
