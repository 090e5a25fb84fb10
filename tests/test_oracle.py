"""Pins the oracle: (1) an independent conformant decoder (FFmpeg's h264 in the bundled libavcodec)
must reproduce the golden model's reconstruction bit-exactly -- north star correctness part 3;
(2) committed bitstream hashes (tests/golden/stream_hashes.json, made by tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

import avdec
from common import content, make_clip, oracle_encode_clip

HERE = os.path.dirname(os.path.abspath(__file__))

def needs_decoder(fn):
    """The independent decoder is the only judge of the (common-mode) standard tables and of the slice data the reference
    cannot pin: a machine without it FAILS these tests instead of skipping them."""
    return fn


def test_independent_decoder_is_present():
    assert avdec.available(), "the bundled libavcodec (opencv-python-headless) is not loadable: the oracle cannot be pinned"



def roundtrip(oracle, kind, w, h, n, fmt=0, **cfg):
    clip = make_clip(kind, w, h, n, fmt)
    stream, sizes, recs = oracle_encode_clip(clip, w, h, fmt, keep_recon=True, **cfg)
    dec = avdec.decode(stream)
    assert len(dec) == n
    for i, (r, d) in enumerate(zip(recs, dec)):
        for p in range(3):
            assert np.array_equal(r[p], d[p]), "frame %d plane %d: decoder != encoder reconstruction" % (i, p)
    return stream


@needs_decoder
@pytest.mark.parametrize("cabac", [0, 1])
@pytest.mark.parametrize("kind,qp", [("synth", 24), ("noise", 1), ("noise", 30), ("static", 36), ("shift", 24),
                                     ("flat", 10), ("synth", 47)])
def test_decode_matches_recon(oracle, kind, qp, cabac):
    roundtrip(oracle, kind, 96, 80, 5, qp=qp, gop=3, cabac=cabac, me_range=8 if kind != "shift" else 16)


@needs_decoder
def test_decode_matches_recon_nv16(oracle):
    roundtrip(oracle, "synth", 64, 48, 3, fmt=1, qp=25, gop=25, cabac=1, me_range=8)


@needs_decoder
def test_decode_matches_recon_padded_size(oracle):
    """src not a multiple of 16 (the README's 854x480 case, scaled down): edge replication to ALIGN16."""
    roundtrip(oracle, "synth", 86, 50, 3, qp=24, gop=25, cabac=1, me_range=8)


@needs_decoder
def test_config1_shape_runs_on_cpu(oracle):
    """BASELINE.json configs[0]: 854x480 NV12, reference defaults (QP 24, GOP 25, CABAC), CPU golden model."""
    stream = roundtrip(oracle, "synth", 854, 480, 2, qp=24, gop=25, cabac=1, me_range=16)
    assert stream.startswith(bytes.fromhex("000000016742".replace("42", "4d")))


def test_golden_stream_hashes(oracle):
    gold = json.load(open(os.path.join(HERE, "golden", "stream_hashes.json")))
    for case in gold["cases"]:
        clip = make_clip(case["kind"], case["w"], case["h"], case["n"], case["fmt"])
        stream, sizes, _ = oracle_encode_clip(clip, case["w"], case["h"], case["fmt"], qp=case["qp"], gop=case["gop"],
                                              cabac=case["cabac"], me_range=case["me"])
        assert sizes == case["sizes"], case
        assert hashlib.sha256(stream).hexdigest() == case["sha256"], case


def test_config_validation_matches_reference_rules(oracle):
    """kernel/cedar.c:744-789: same accept/reject decisions, -EINVAL."""
    import ctypes as C
    L = oracle.lib()

    def rc(**kw):
        base = dict(width=64, height=48)
        base.update(kw)
        relax = base.pop("relax_gop", 0)
        cfg = oracle.make_config(relax_gop=relax, **base)
        h = C.c_void_p()
        r = L.gm_open(C.byref(cfg), C.byref(h))
        if r == 0:
            L.gm_close(h)
        return r

    assert rc() == 0
    assert rc(width=63) == -22 and rc(height=47) == -22          # src not even
    assert rc(dst_width=70) == -22                                # dst not multiple of 16
    assert rc(width=80, dst_width=64) == -22                      # src > dst
    assert rc(qp=0) == -22 and rc(qp=48) == -22 and rc(qp=47) == 0 and rc(qp=1) == 0
    assert rc(fmt=2) == -22
    assert rc(gop=0) == -22 and rc(gop=32) == -22 and rc(gop=31) == 0
    assert rc(gop=60, relax_gop=1) == 0                           # documented extension


def test_gop_structure_and_frame_num(oracle):
    enc = oracle.Encoder(oracle.make_config(64, 48, gop=3))
    kinds = []
    for t in range(7):
        y, c = oracle.synth_frame(64, 48, t)
        enc.encode(y, c)
        kinds.append(enc.frame_is_i())
    assert kinds == [True, False, False, True, False, False, True]


def test_skip_macroblocks_appear_on_static_content(oracle):
    enc = oracle.Encoder(oracle.make_config(96, 80, qp=36, gop=5, cabac=0, me_range=8))
    for t in range(2):
        y, c = content("static", 96, 80, t)
        enc.encode(y, c)
    assert (enc.mbs()["type"] == 2).sum() >= 15


@needs_decoder
@pytest.mark.parametrize("cabac", [0, 1])
@pytest.mark.parametrize("kind,qp", [("synth", 24), ("noise", 12), ("shift", 36)])
def test_intra4x4_option_decodes_bit_exactly(oracle, kind, qp, cabac):
    """Opt-in Intra4x4 (oracle only so far): nine prediction modes, I_NxN syntax for both
    entropy coders.  Off by default: with the SAD-based I4x4/I16x16 decision it costs bits."""
    clip = make_clip(kind, 96, 80, 3)
    stream, sizes, recs = oracle_encode_clip(clip, 96, 80, keep_recon=True, qp=qp, gop=2, cabac=cabac, me_range=8, intra4x4=1)
    dec = avdec.decode(stream)
    assert len(dec) == 3
    for r, d in zip(recs, dec):
        for p in range(3):
            assert np.array_equal(r[p], d[p])


@needs_decoder
@pytest.mark.parametrize("cabac", [0, 1])
@pytest.mark.parametrize("kind,qp,rows,i4", [("synth", 24, 1, 0), ("noise", 12, 2, 0), ("shift", 36, 3, 0), ("static", 30, 2, 0),
                                             ("synth", 26, 2, 1)])
def test_slice_rows_option_decodes_bit_exactly(oracle, kind, qp, rows, i4, cabac):
    """slice_rows extension (north star: "MB rows per slice configurable"): every slice is its own NAL with
    first_mb_in_slice != 0; neighbours in other slices are unavailable for intra / MV prediction and all
    entropy contexts, deblocking still runs across slice edges.  The independent decoder must agree."""
    w, h, n = 96, 80, 4
    clip = make_clip(kind, w, h, n)
    stream, sizes, recs = oracle_encode_clip(clip, w, h, keep_recon=True, qp=qp, gop=3, cabac=cabac, me_range=8,
                                             slice_rows=rows, intra4x4=i4)
    nslices = -(-(h // 16) // rows)
    assert stream.count(b"\x00\x00\x00\x01\x65") + stream.count(b"\x00\x00\x00\x01\x41") >= n * nslices
    dec = avdec.decode(stream)
    assert len(dec) == n
    for r, d in zip(recs, dec):
        for p in range(3):
            assert np.array_equal(r[p], d[p])


def test_slice_rows_zero_and_full_height_are_the_reference_layout(oracle):
    clip = make_clip("synth", 64, 48, 3)
    a, _, _ = oracle_encode_clip(clip, 64, 48, qp=24, gop=2, cabac=1, me_range=8)
    b, _, _ = oracle_encode_clip(clip, 64, 48, qp=24, gop=2, cabac=1, me_range=8, slice_rows=3)
    c, _, _ = oracle_encode_clip(clip, 64, 48, qp=24, gop=2, cabac=1, me_range=8, slice_rows=99)
    assert a == b == c


@needs_decoder
def test_sps_crop_makes_the_decoder_output_the_source_size(oracle):
    w, h, n = 100, 50, 3
    clip = make_clip("synth", w, h, n)
    stream, _, recs = oracle_encode_clip(clip, w, h, keep_recon=True, qp=24, gop=2, cabac=1, me_range=8, sps_crop=1,
                                         auto_level=1)
    dec = avdec.decode(stream)
    assert len(dec) == n
    for r, d in zip(recs, dec):
        assert d[0].shape == (h, w) and d[1].shape == (h // 2, w // 2)
        assert np.array_equal(r[0][:h, :w], d[0]) and np.array_equal(r[1][:h // 2, :w // 2], d[1])


# ---- the golden model's own decoder (oracle/h264_decoder.c) ---------------------------------------
@pytest.mark.parametrize("cabac", [0, 1])
@pytest.mark.parametrize("kind,qp,rows,i4", [("synth", 24, 0, 0), ("noise", 12, 0, 0), ("static", 36, 0, 0), ("shift", 30, 2, 0),
                                             ("synth", 26, 1, 1), ("noise", 1, 0, 1), ("flat", 47, 3, 0)])
def test_golden_decoder_reproduces_the_encoder_reconstruction(oracle, kind, qp, rows, i4, cabac):
    """North star part 3 without libavcodec: parsing the stream from the decoding side of the standard (CAVLC by table
    matching, the CABAC decoding engine) and reconstructing gives exactly the encoder's reference pictures."""
    w, h, n = 96, 80, 5
    clip = make_clip(kind, w, h, n)
    stream, _, recs = oracle_encode_clip(clip, w, h, keep_recon=True, qp=qp, gop=3, cabac=cabac, me_range=8,
                                         slice_rows=rows, intra4x4=i4)
    dec = oracle.golden_decode(stream)
    assert len(dec) == n
    for r, d in zip(recs, dec):
        for p in range(3):
            assert np.array_equal(r[p], d[p])


@needs_decoder
def test_golden_decoder_agrees_with_libavcodec(oracle):
    w, h, n = 100, 50, 4
    clip = make_clip("synth", w, h, n)
    stream, _, _ = oracle_encode_clip(clip, w, h, qp=22, gop=2, cabac=1, me_range=16, slice_rows=2, sps_crop=1,
                                      repeat_headers=1)
    a, b = avdec.decode(stream), oracle.golden_decode(stream)
    assert len(a) == len(b) == n
    for x, y in zip(a, b):
        for p in range(3):
            assert x[p].shape == y[p].shape and np.array_equal(x[p], y[p])


def test_golden_decoder_rejects_what_it_does_not_cover(oracle):
    clip = make_clip("synth", 64, 48, 2)
    stream, _, _ = oracle_encode_clip(clip, 64, 48, qp=24, gop=2, cabac=0, me_range=8)
    with pytest.raises(ValueError):
        oracle.golden_decode(stream[stream.index(b"\x00\x00\x00\x01\x65"):])  # slice before SPS / PPS


@needs_decoder
@pytest.mark.parametrize("cabac", [0, 1])
@pytest.mark.parametrize("kind,qp,rows", [("noise", 30, 0), ("noise", 12, 3), ("shift", 36, 2)])
def test_p_intra_option_decodes_bit_exactly(oracle, kind, qp, rows, cabac):
    """p_intra extension: Intra16x16 macroblocks inside P slices (mb_type 5 + ..., intra prediction from inter
    neighbours, mixed-edge deblocking) -- libavcodec and the golden decoder both reproduce the reconstruction."""
    clip = make_clip(kind, 96, 80, 5)
    stream, _, recs = oracle_encode_clip(clip, 96, 80, keep_recon=True, qp=qp, gop=4, cabac=cabac, me_range=8, slice_rows=rows,
                                         p_intra=1)
    for dec in (avdec.decode(stream), oracle.golden_decode(stream)):
        assert len(dec) == 5
        for r, d in zip(recs, dec):
            for p in range(3):
                assert np.array_equal(r[p], d[p])
