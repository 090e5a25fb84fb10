"""Experiment: k encoder handles on one GPU, driven from k host threads, against one handle over the same frames.
Prints frames/s (wall clock around issue + synchronize, warm)."""
import sys, time, threading
sys.path.insert(0, ".")
import torch
import cedarx_h264_encoder_b200 as cx
from cedarx_h264_encoder_b200 import api, synth

w, h, gop, qp, me = 1920, 1080, 60, 25, 16


def make(n, lanes):
    cfg = api.make_config(w, h, qp=qp, gop=gop, cabac=1, me_range=me, max_clip_frames=n, gops_in_flight=lanes)
    enc = cx.Encoder(cfg)
    staging = torch.from_numpy(enc.clip_input(n))
    for i in range(0, n, 20):
        part = synth.synth_clip(w, h, list(range(i, min(n, i + 20))), 0, device="cuda")
        staging[i:i + len(part)].copy_(part)
    torch.cuda.synchronize()
    enc.clip_upload(n)
    return enc


def run(encs, n, iters=4):
    def work(e):
        for _ in range(iters):
            e.clip_encode(n, 0)
    for e in encs:
        e.clip_encode(n, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(e,)) for e in encs]
    [t.start() for t in th]
    [t.join() for t in th]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return len(encs) * n * iters / dt


for k, n, lanes in [(1, 600, 10), (2, 300, 5), (3, 180, 3), (5, 120, 2), (2, 600, 10), (3, 600, 10)]:
    encs = [make(n, lanes) for _ in range(k)]
    print("handles %d x %d frames (lanes %d): %.0f frames/s" % (k, n, lanes, run(encs, n)), flush=True)
    for e in encs:
        e.close()
