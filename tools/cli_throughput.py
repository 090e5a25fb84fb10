#!/usr/bin/env python3
"""CLI end to end (SURVEY 8d ii): a synthetic 1080p clip in a file (page cache), through cedarx_h264_encoder_b200/h264enc:
frame at a time (the reference's synchronous flow), --queue-gops (the same loop on a queued handle), --batch-gops (reader
thread + pipeline workers + ordered writer) and --gpus N when the box has several GPUs.  Prints one JSON line with
frames/s of each (process start, pinned allocations and file I/O included) and checks that all outputs are identical.

    python tools/cli_throughput.py [frames] [out.json] [modes, comma separated]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cedarx_h264_encoder_b200 import synth  # noqa: E402

w, h, n, gop = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 600, 60
raw = "/tmp/clip_1080p.nv12"
with open(raw, "wb") as f:
    for i in range(0, n, 20):
        f.write(synth.synth_clip(w, h, list(range(i, min(n, i + 20))), 0, device="cuda").cpu().numpy().tobytes())
cli = os.path.join(ROOT, "cedarx_h264_encoder_b200", "h264enc")
modes = [("frame_at_a_time", []), ("queue_gops_2", ["--queue-gops", "2"]), ("batch_gops_2_handles_2", ["--batch-gops", "2"]),
         ("batch_gops_5_handles_2", ["--batch-gops", "5"]), ("batch_gops_2_handles_3", ["--batch-gops", "2", "--handles", "3"])]
modes += [("batch_gops_2_reader_1", ["--batch-gops", "2", "--reader-threads", "1"]),
          ("batch_gops_2_handles_3_reader_8", ["--batch-gops", "2", "--handles", "3", "--reader-threads", "8"])]
ngpu = torch.cuda.device_count()
if ngpu > 1:
    modes.append(("gpus_%d_batch_gops_2" % ngpu, ["--gpus", str(ngpu), "--batch-gops", "2"]))
if len(sys.argv) > 3:
    modes = [m for m in modes if m[0] in sys.argv[3].split(",")]
res, outs = {}, {}
for name, extra in modes:
    out = "/tmp/out_%s.264" % name
    best, timing = None, ""
    for rep in range(2):  # second run: input in the page cache, driver warm
        t = time.time()
        r = subprocess.run([cli, raw, str(w), str(h), out, "--qp", "25", "--gop", str(gop), "--stats"] + extra, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        dt = time.time() - t
        if best is None or dt < best:
            best = dt
            timing = [ln for ln in r.stderr.splitlines() if ln.startswith("timing:")][-1]
    # "timing: open A s, stream B s (C frames/s), close D s": the stream phase is reading, encoding and writing, overlapped
    parts = timing.replace("(", "").split()
    res[name] = {"frames_per_s": round(n / best, 1), "wall_s": round(best, 3), "open_s": float(parts[2]),
                 "stream_s": float(parts[5]), "stream_frames_per_s": float(parts[7]), "close_s": float(parts[10])}
    outs[name] = open(out, "rb").read()
    print("%-26s %8.1f frames/s (%.2f s wall) | %s" % (name, n / best, best, timing), file=sys.stderr)
first = next(iter(outs.values()))
same = all(v == first for v in outs.values())
line = {"tool": "cli_throughput", "clip": "%dx%d nv12, %d frames, GOP %d, QP 25, from a file" % (w, h, n, gop), "gpus": ngpu,
        "modes": res, "outputs_identical": same, "bytes": len(first)}
print(json.dumps(line))
if len(sys.argv) > 2:
    json.dump(line, open(sys.argv[2], "w"), indent=1)
assert same, "outputs differ"
