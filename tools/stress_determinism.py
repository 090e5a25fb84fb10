#!/usr/bin/env python3
"""Stress check for the wavefront / multi-stream synchronisation (the stand-in for compute-sanitizer's racecheck, which
is closed on this pool): the 600-frame 1080p clip must encode to the same bytes on every run, for every number of GOPs
in flight, with three handles hammering the GPU from three host threads at once, and libavcodec must decode it.

    python tools/stress_determinism.py [out.json]      (env: N frames, ROUNDS, CABAC)"""
import hashlib, json, os, sys, threading
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cedarx_h264_encoder_b200 as cx
from cedarx_h264_encoder_b200 import api, synth
w, h, n, gop = 1920, 1088, int(os.environ.get("N", 600)), 60
rounds = int(os.environ.get("ROUNDS", 8))
report = {"tool": "stress_determinism", "clip": "%dx%d, %d frames, GOP %d, QP 25" % (w, h, n, gop), "cases": []}
clip = torch.empty((n, w * h * 3 // 2), dtype=torch.uint8)
for i in range(0, n, 20):
    clip[i:i + 20].copy_(synth.synth_clip(w, h, list(range(i, min(n, i + 20))), 0, device="cuda"))
ref = {}
for cabac in (1, 0):
    # (a) one handle at a time, different numbers of GOPs in flight
    for lanes, reps in ((0, 3), (1, 1), (3, 1), (7, 1)):
        enc = cx.Encoder(api.make_config(w, h, qp=25, gop=gop, cabac=cabac, max_clip_frames=n, gops_in_flight=lanes))
        torch.from_numpy(enc.clip_input(n)).copy_(clip)
        hashes = []
        for r in range(reps):
            enc.clip_upload(n); enc.clip_encode(n, 0)
            data, sizes = enc.clip_download(n)
            hashes.append(hashlib.sha256(data.tobytes()).hexdigest())
            if cabac not in ref:
                ref[cabac] = hashes[-1]
                import avdec
                dec = avdec.decode(data[:int(sizes[:3].sum())].tobytes())
                assert len(dec) == 3
        enc.close()
        ok = all(x == ref[cabac] for x in hashes)
        report["cases"].append({"entropy": "cabac" if cabac else "cavlc", "handles": 1, "gops_in_flight": lanes or "auto",
                                "runs": reps, "identical": ok})
        print(report["cases"][-1], flush=True)
        assert ok, "NON-DETERMINISTIC OUTPUT"
    # (b) three handles on three host threads, all in flight together, several rounds each
    encs = [cx.Encoder(api.make_config(w, h, qp=25, gop=gop, cabac=cabac, max_clip_frames=n)) for _ in range(3)]
    for e in encs:
        torch.from_numpy(e.clip_input(n)).copy_(clip)
    got = [[] for _ in encs]

    def work(i):
        for _ in range(rounds):
            encs[i].clip_upload(n); encs[i].clip_encode(n, 0)
            got[i].append(hashlib.sha256(encs[i].clip_download(n)[0].tobytes()).hexdigest())
    th = [threading.Thread(target=work, args=(i,)) for i in range(3)]
    [t.start() for t in th]; [t.join() for t in th]
    for e in encs:
        e.close()
    ok = all(x == ref[cabac] for g in got for x in g) and all(len(g) == rounds for g in got)
    report["cases"].append({"entropy": "cabac" if cabac else "cavlc", "handles": 3, "threads": 3, "runs": 3 * rounds, "identical": ok})
    print(report["cases"][-1], flush=True)
    assert ok, "NON-DETERMINISTIC OUTPUT"
report["sha256"] = {("cabac" if k else "cavlc"): v for k, v in ref.items()}
report["all_identical"] = True
print("STRESS OK")
if len(sys.argv) > 1:
    json.dump(report, open(sys.argv[1], "w"), indent=1)
