"""The product's entropy logic (csrc/entropy.cuh: CAVLC count/scan/scatter, CABAC binarisation +
byte-wise arithmetic coder, order-independent emulation prevention) compiled for the host and driven
like the CUDA kernels drive it, against the oracle's bitstream -- no GPU needed."""
import numpy as np
import pytest

import avdec
from common import content

MBI = np.dtype([("type", "u1"), ("i16_mode", "u1"), ("chroma_mode", "u1"), ("cbp", "u1"), ("mv", "<i2", (2,)),
                ("mvd", "<i2", (2,)), ("pad", "<u4")])


def to_product_layout(mbs):
    n = len(mbs)
    mbi = np.zeros(n, MBI)
    for k in ("type", "i16_mode", "chroma_mode", "cbp", "mv", "mvd"):
        mbi[k] = mbs[k]
    nnz = np.zeros((n, 32), np.uint8)
    nnz[:, :27] = mbs["nnz"]
    coef = np.ascontiguousarray(mbs["coef"]).reshape(n, 26 * 16)
    return mbi, nnz, coef, np.ascontiguousarray(mbs["i4_mode"])


@pytest.mark.parametrize("cabac", [0, 1])
@pytest.mark.parametrize("kind,qp,rows,i4", [("synth", 24, 0, 0), ("noise", 1, 0, 0), ("noise", 12, 0, 0), ("noise", 40, 0, 0),
                                             ("static", 36, 0, 0), ("synth", 24, 1, 0), ("static", 36, 3, 0), ("noise", 30, 2, 0),
                                             ("synth", 24, 0, 1), ("noise", 12, 2, 1), ("shift", 36, 0, 1)])
def test_entropy_logic_matches_oracle(oracle, harness, kind, qp, rows, i4, cabac):
    """rows = slice_rows: every slice NAL of the picture goes through the product logic on its own."""
    w, h, gop = 96, 80, 3
    mbw, mbh = w // 16, h // 16
    srows = rows if rows else mbh
    nslices = -(-mbh // srows)
    enc = oracle.Encoder(oracle.make_config(w, h, qp=qp, gop=gop, cabac=cabac, me_range=8, slice_rows=rows, intra4x4=i4))
    for t in range(5):
        y, c = content(kind, w, h, t)
        payloads = [nal[5:] for ty, nal in avdec.split_nals(enc.encode(y, c)) if ty in (1, 5)]
        assert len(payloads) == nslices
        mbi, nnz, coef, i4m = to_product_layout(enc.mbs())
        fi = int(enc.frame_is_i())
        for k, payload in enumerate(payloads):
            bits = oracle.slice_header_bits(fi, t % gop, cabac, k * srows * mbw)
            out = np.zeros(len(payload) * 2 + 4096, np.uint8)
            if cabac:
                n = harness.hh_cabac_slice(mbi.ctypes.data, nnz.ctypes.data, coef.ctypes.data, mbw, mbh, srows, k, fi, qp,
                                           int(bits, 2), len(bits), out.ctypes.data, out.size, [62, 7, 1000][t % 3], i4m.ctypes.data)
            else:
                n = harness.hh_cavlc_slice(mbi.ctypes.data, nnz.ctypes.data, coef.ctypes.data, mbw, mbh, srows, k, fi,
                                           int(bits, 2), len(bits), out.ctypes.data, out.size, i4m.ctypes.data)
            assert n > 0
            esc = np.zeros(n * 2 + 16, np.uint8)
            m = harness.hh_epb(out.ctypes.data, n, esc.ctypes.data, esc.size)
            assert esc[:m].tobytes() == payload, "frame %d (%s) slice %d" % (t, "I" if fi else "P", k)


def test_parallel_emulation_prevention_rule(harness):
    """epb_needed(byte, zero_run) must reproduce the sequential rule on adversarial zero runs."""
    rng = np.random.default_rng(3)
    for trial in range(200):
        n = int(rng.integers(1, 200))
        data = rng.choice(np.array([0, 0, 0, 1, 2, 3, 4, 255], np.uint8), size=n)
        want, zeros = bytearray(), 0
        for b in data.tolist():
            if zeros >= 2 and b <= 3:
                want.append(3)
                zeros = 0
            want.append(b)
            zeros = zeros + 1 if b == 0 else 0
        out = np.zeros(2 * n + 8, np.uint8)
        m = harness.hh_epb(data.ctypes.data, n, out.ctypes.data, out.size)
        assert out[:m].tobytes() == bytes(want)


def test_mps_path_renormalises_by_at_most_one_bit(harness):
    """CabacRange::step computes the MPS shift as (rm >> 8) ^ 1; valid iff range - rangeLPS >= 128 always."""
    assert harness.hh_mps_renorm_at_most_one() == 1
