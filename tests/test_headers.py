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
