#!/usr/bin/env python3
"""CLI end to end (SURVEY 8d ii): a synthetic 1080p clip in a file, through cedarx_h264_encoder_b200/h264enc, frame at a
time (the reference's flow) and GOP-parallel (--batch-gops); prints frames/s of both and checks the outputs are equal."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cedarx_h264_encoder_b200 import synth  # noqa: E402

w, h, n, gop = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 600, 60
raw = "/tmp/clip_1080p.nv12"
with open(raw, "wb") as f:
    for i in range(0, n, 20):
        f.write(synth.synth_clip(w, h, list(range(i, min(n, i + 20))), 0, device="cuda").cpu().numpy().tobytes())
cli = os.path.join(ROOT, "cedarx_h264_encoder_b200", "h264enc")
res = {}
for name, extra in (("frame_at_a_time", []), ("batch_gops_10", ["--batch-gops", "10"])):
    out = "/tmp/out_%s.264" % name
    t = time.time()
    subprocess.run([cli, raw, str(w), str(h), out, "--qp", "25", "--gop", str(gop)] + extra, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dt = time.time() - t
    res[name] = (n / dt, open(out, "rb").read())
    print("%-16s %7.1f frames/s (%.2f s wall, process start and file I/O included)" % (name, n / dt, dt))
assert res["frame_at_a_time"][1] == res["batch_gops_10"][1], "outputs differ"
print("outputs identical: %d bytes" % len(res["batch_gops_10"][1]))
