#!/usr/bin/env python3
"""Stage-by-stage comparison of the CUDA encoder with the CPU golden model (development aid;
the pytest version of these checks lives in tests/test_gpu_parity.py).  Run on a GPU box:
    python tests/gpu_check.py [quick|full]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cedarx_h264_encoder_b200 as cx  # noqa: E402
from cedarx_h264_encoder_b200 import api  # noqa: E402
import oracle_lib as O  # noqa: E402


def content(kind, w, h, t):
    if kind == "synth":
        return O.synth_frame(w, h, t)
    if kind == "noise":
        r = np.random.default_rng(100 + t)
        return r.integers(0, 256, (h, w), dtype=np.uint8), r.integers(0, 256, (h // 2, w), dtype=np.uint8)
    if kind == "static":
        y, c = O.synth_frame(w, h, 0)
        if t % 3 == 2:
            y = y.copy()
            y[16:48, 32:80] = O.synth_frame(w, h, t)[0][16:48, 32:80]
        return y, c
    raise ValueError(kind)


def first_diff(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    d = np.argwhere(a != b)
    return d[0].tolist() if len(d) else None


def check_frame_mode(kind, w, h, n, qp, gop, cabac, me):
    ocfg = O.make_config(w, h, qp=qp, gop=gop, cabac=cabac, me_range=me)
    gcfg = api.make_config(w, h, qp=qp, gop=gop, cabac=cabac, me_range=me)
    oe = O.Encoder(ocfg)
    ge = cx.Encoder(gcfg)
    ok = True
    for t in range(n):
        y, c = content(kind, w, h, t)
        ob = oe.encode(y, c)
        gb = ge.encode(y, c)
        msgs = []
        osrc, gsrc = oe.source(), ge.debug_planes(0)
        for p in range(3):
            if not np.array_equal(osrc[p], gsrc[p]):
                msgs.append("src plane %d differs at %s" % (p, first_diff(osrc[p], gsrc[p])))
        ombs = oe.mbs()
        gmbi, gnnz, gcoef = ge.debug_syntax()
        for k in ("type", "i16_mode", "chroma_mode", "cbp", "mv", "mvd"):
            if not np.array_equal(ombs[k], gmbi[k]):
                i = first_diff(ombs[k], gmbi[k])
                msgs.append("mb.%s differs at %s: oracle %s gpu %s (n=%d)" % (
                    k, i, ombs[k][i[0]].tolist(), gmbi[k][i[0]].tolist(), int((ombs[k] != gmbi[k]).sum())))
        if not np.array_equal(ombs["nnz"], gnnz[:, :27]):
            i = first_diff(ombs["nnz"], gnnz[:, :27])
            msgs.append("nnz differs at %s: oracle %s gpu %s" % (i, ombs["nnz"][i[0]].tolist(), gnnz[i[0], :27].tolist()))
        if not np.array_equal(ombs["coef"], gcoef):
            i = first_diff(ombs["coef"], gcoef)
            msgs.append("coef differs at %s: oracle %s gpu %s (n=%d)" % (
                i, ombs["coef"][i[0], i[1]].tolist(), gcoef[i[0], i[1]].tolist(), int((ombs["coef"] != gcoef).sum())))
        ounf, gunf = oe.recon_unfiltered(), ge.debug_planes(1)
        for p in range(3):
            if not np.array_equal(ounf[p], gunf[p]):
                msgs.append("unfiltered recon plane %d differs at %s (n=%d)" % (
                    p, first_diff(ounf[p], gunf[p]), int((ounf[p] != gunf[p]).sum())))
        orec, grec = oe.recon(), ge.debug_planes(2)
        for p in range(3):
            if not np.array_equal(orec[p], grec[p]):
                msgs.append("deblocked recon plane %d differs at %s (n=%d)" % (
                    p, first_diff(orec[p], grec[p]), int((orec[p] != grec[p]).sum())))
        if ob != gb:
            k = next((i for i in range(min(len(ob), len(gb))) if ob[i] != gb[i]), min(len(ob), len(gb)))
            msgs.append("bitstream differs: len oracle %d gpu %d first diff %d: %s vs %s" % (
                len(ob), len(gb), k, ob[max(0, k - 4):k + 8].hex(), gb[max(0, k - 4):k + 8].hex()))
        if abs(oe.sse_y() - ge.sse_y(1)[0]) > 0.5:
            msgs.append("sse differs: %f vs %f" % (oe.sse_y(), ge.sse_y(1)[0]))
        if msgs:
            ok = False
            print("  frame %d (%s):" % (t, "I" if oe.frame_is_i() else "P"))
            for m in msgs[:12]:
                print("     " + m)
    print("%s frame-mode %s %dx%d n=%d qp=%d gop=%d cabac=%d me=%d" % ("OK " if ok else "BAD", kind, w, h, n, qp, gop, cabac, me))
    ge.close()
    oe.close()
    return ok


def make_clip(kind, w, h, n):
    fb = w * h * 3 // 2
    clip = np.empty((n, fb), np.uint8)
    for t in range(n):
        y, c = content(kind, w, h, t)
        clip[t, :w * h] = y.reshape(-1)
        clip[t, w * h:] = c.reshape(-1)
    return clip


def check_clip_mode(kind, w, h, n, qp, gop, cabac, me, lanes=0):
    clip = make_clip(kind, w, h, n)
    oe = O.Encoder(O.make_config(w, h, qp=qp, gop=gop, cabac=cabac, me_range=me))
    ref = b""
    osz = []
    for t in range(n):
        b = oe.encode(clip[t, :w * h].reshape(h, w), clip[t, w * h:].reshape(h // 2, w))
        ref += b
        osz.append(len(b))
    ge = cx.Encoder(api.make_config(w, h, qp=qp, gop=gop, cabac=cabac, me_range=me, max_clip_frames=n, gops_in_flight=lanes))
    t0 = time.time()
    got, sizes = ge.encode_clip(clip)
    dt = time.time() - t0
    ok = got == ref and sizes.tolist() == osz
    print("%s clip-mode %s %dx%d n=%d qp=%d gop=%d cabac=%d me=%d lanes=%d: %d bytes (oracle %d), %.1f ms" % (
        "OK " if ok else "BAD", kind, w, h, n, qp, gop, cabac, me, lanes, len(got), len(ref), dt * 1e3))
    if not ok:
        bad = [i for i in range(n) if i >= len(sizes) or sizes[i] != osz[i]]
        print("     frames with different sizes:", bad[:10], "gpu", sizes[:8].tolist(), "oracle", osz[:8])
    ge.close()
    oe.close()
    return ok


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    print(cx.load_library().cedar_b200_version().decode())
    ok = True
    ok &= check_frame_mode("synth", 96, 80, 4, 24, 3, 0, 8)
    ok &= check_frame_mode("synth", 96, 80, 4, 24, 3, 1, 8)
    ok &= check_frame_mode("noise", 96, 80, 3, 30, 3, 1, 8)
    ok &= check_frame_mode("static", 96, 80, 5, 30, 5, 0, 8)
    ok &= check_frame_mode("synth", 176, 144, 3, 25, 25, 1, 16)
    ok &= check_clip_mode("synth", 96, 80, 12, 24, 4, 0, 8)
    ok &= check_clip_mode("synth", 96, 80, 12, 24, 4, 1, 8)
    ok &= check_clip_mode("synth", 96, 80, 11, 24, 4, 1, 8, lanes=2)
    if mode == "full":
        for cabac in (0, 1):
            for qp in (1, 12, 36, 47):
                ok &= check_frame_mode("noise", 96, 80, 3, qp, 3, cabac, 8)
                ok &= check_frame_mode("synth", 96, 80, 3, qp, 3, cabac, 8)
        ok &= check_frame_mode("synth", 854, 480, 3, 24, 25, 1, 16)
        ok &= check_clip_mode("synth", 352, 288, 20, 25, 5, 1, 16)
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
