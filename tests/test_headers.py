"""Header bytes must be byte-identical to what the reference's writer emits (north star, part 2).
Both the oracle's restatement and the product's host-C writer are checked against the same
known-answer vectors (tests/golden/headers.json) and against each other over a sweep."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "headers.json")))


def _hex(b):
    return b.hex(" ")


@pytest.mark.parametrize("dims", sorted(GOLD["sps"]))
def test_sps_known_answers(oracle, product_lib, dims):
    from cedarx_h264_encoder_b200 import api
    w, h = map(int, dims.split("x"))
    assert _hex(oracle.write_sps(oracle.make_config(w, h))) == GOLD["sps"][dims]
    assert _hex(api.write_sps(api.make_config(w, h))) == GOLD["sps"][dims]


@pytest.mark.parametrize("key", sorted(GOLD["pps"]))
def test_pps_known_answers(oracle, product_lib, key):
    from cedarx_h264_encoder_b200 import api
    qp = int(key[2:4])
    cabac = int(key.endswith("cabac"))
    assert _hex(oracle.write_pps(oracle.make_config(64, 64, qp=qp, cabac=cabac))) == GOLD["pps"][key]
    assert _hex(api.write_pps(api.make_config(64, 64, qp=qp, cabac=cabac))) == GOLD["pps"][key]


def test_sps_trailing_bits_quirk(oracle):
    """cedar.c:883-890: with 7 bits used in the last byte the stop bit is followed by a whole 0x00."""
    assert oracle.write_sps(oracle.make_config(1920, 1088)).endswith(b"\x91\x00")
    assert not oracle.write_sps(oracle.make_config(1280, 720)).endswith(b"\x00")


def test_slice_header_bits(oracle, product_lib):
    from cedarx_h264_encoder_b200 import api
    s = GOLD["slice"]
    for f in (oracle.slice_header_bits, api.slice_header_bits):
        assert f(1, 0, 1) == s["i"] and f(1, 0, 0) == s["i"]
        for fpc in (1, 7, 15, 16, 24, 59):
            ffff = format(fpc & 15, "04b")
            assert f(0, fpc, 1) == s["p_cabac_prefix"] + ffff + s["p_cabac_suffix"]
            assert f(0, fpc, 0) == s["p_cavlc_prefix"] + ffff + s["p_cavlc_suffix"]


def test_product_headers_equal_oracle_sweep(oracle, product_lib):
    from cedarx_h264_encoder_b200 import api
    for w in range(16, 4097, 208):
        for h in (16, 480, 1088, 2160):
            assert api.write_sps(api.make_config(w, h)) == oracle.write_sps(oracle.make_config(w, h))
    for qp in range(1, 48):
        for cabac in (0, 1):
            assert api.write_pps(api.make_config(64, 64, qp=qp, cabac=cabac)) == \
                oracle.write_pps(oracle.make_config(64, 64, qp=qp, cabac=cabac))


def test_first_frame_carries_sps_pps_once(oracle):
    """cedar.c:1058-1061: parameter sets only before the very first frame, not at later IDRs."""
    import avdec
    enc = oracle.Encoder(oracle.make_config(64, 48, gop=2))
    types = []
    for t in range(5):
        y, c = oracle.synth_frame(64, 48, t)
        types.append([n[0] for n in avdec.split_nals(enc.encode(y, c))])
    assert types == [[7, 8, 5], [1], [5], [1], [5]]


def _spec_sps(profile, level, wmb, hmb, crop_right, crop_bottom):
    """SPS written straight from the syntax table of H.264 7.3.2.1.1 (independent of both writers)."""
    bits = ""

    def u(v, n):
        nonlocal bits
        bits += format(v, "0%db" % n)

    def ue(v):
        nonlocal bits
        s = format(v + 1, "b")
        bits += "0" * (len(s) - 1) + s

    u(profile, 8); u(0, 8); u(level, 8)
    ue(0); ue(0); ue(2); ue(1); u(0, 1); ue(wmb - 1); ue(hmb - 1); u(1, 1); u(0, 1)
    if crop_right or crop_bottom:
        u(1, 1); ue(0); ue(crop_right); ue(0); ue(crop_bottom)
    else:
        u(0, 1)
    u(0, 1)
    bits += "1"                       # rbsp_stop_one_bit
    extra = len(bits) % 8 == 0        # cedar.c:883-890 quirk: stop bit in the last position is followed by a whole 0x00
    bits += "0" * (-len(bits) % 8)
    body = int(bits, 2).to_bytes(len(bits) // 8, "big") + (b"\x00" if extra else b"")
    out, zeros = bytearray(b"\x00\x00\x00\x01\x67"), 0
    for b in body:
        if zeros >= 2 and b <= 3:
            out.append(3)
            zeros = 0
        out.append(b)
        zeros = zeros + 1 if b == 0 else 0
    return bytes(out)


@pytest.mark.parametrize("w,h", [(1920, 1080), (854, 480), (100, 50), (3840, 2160), (1280, 720), (64, 48)])
def test_sps_crop_and_auto_level_extensions(oracle, product_lib, w, h):
    """sps_crop / auto_level: off by default (reference bytes), on = cropping syntax in the field
    order of the reference's dead branch (cedar.c:924-929) with offsets in crop units, level from table A-1."""
    from cedarx_h264_encoder_b200 import api
    wmb, hmb = (w + 15) // 16, (h + 15) // 16
    levels = {(1920, 1080): 40, (854, 480): 22, (100, 50): 10, (3840, 2160): 51, (1280, 720): 31, (64, 48): 10}
    plain = _spec_sps(77, 41, wmb, hmb, 0, 0)
    assert api.write_sps(api.make_config(w, h)) == oracle.write_sps(oracle.make_config(w, h)) == plain
    want = _spec_sps(77, levels[(w, h)], wmb, hmb, (wmb * 16 - w) // 2, (hmb * 16 - h) // 2)
    assert api.write_sps(api.make_config(w, h, sps_crop=1, auto_level=1)) == want
    assert oracle.write_sps(oracle.make_config(w, h, sps_crop=1, auto_level=1)) == want
