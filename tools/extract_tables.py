#!/usr/bin/env python3
"""Generate the H.264 standard-table headers used by the oracle and by the CUDA product.

There is no copy of ITU-T H.264 offline.  The tables below are *standard data* (CABAC
context initialisation, rangeTabLPS, CAVLC VLC tables, deblocking alpha/beta/tc0, ...).
They are located by byte signature inside the libavcodec that ships with
opencv-python-headless, sanity-checked against a few values known
from the standard, and written out as plain C arrays.  The generated headers are committed,
so nothing at build, test or run time depends on this script or on libavcodec.

Usage: python tools/extract_tables.py   (rewrites cedarx_h264_encoder_b200/csrc/h264_tables.h, the one copy in the
       tree; oracle/h264_tables.h only includes it)
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_lib():
    import cv2  # noqa: F401  (locates site-packages)
    base = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    c = glob.glob(os.path.join(base, "libavcodec-*.so*"))
    if not c:
        sys.exit("libavcodec not found")
    return c[0]


def main():
    d = open(find_lib(), "rb").read()

    def at(sig, nth=0):
        b = bytes.fromhex(sig.replace(" ", ""))
        i = d.find(b)
        for _ in range(nth):
            i = d.find(b, i + 1)
        if i < 0:
            sys.exit("signature not found: " + sig)
        return i

    def u8(o, n):
        return np.frombuffer(d[o:o + n], dtype=np.uint8).astype(np.int32)

    def i8(o, n):
        return np.frombuffer(d[o:o + n], dtype=np.int8).astype(np.int32)

    out = {}
    # --- CABAC context init: 4 tables (PB idc0, idc1, idc2, I), [1024][2] int8 (m, n)
    o = at("14 f1 02 36 03 4a 14 f1 02 36 03 4a e4 7f e9 68")
    pb0 = i8(o, 2048).reshape(1024, 2)[:460]
    it = i8(o + 3 * 0x800, 2048).reshape(1024, 2)[:460]
    assert pb0[0].tolist() == [20, -15] and pb0[11].tolist() == [23, 33]
    assert it[60].tolist() == [0, 41] and it[276 - 0 if False else 70].tolist() == [0, 11]
    out["cabac_init_I"] = it
    out["cabac_init_P0"] = pb0

    # --- rangeTabLPS stored [q][state][2 copies]; followed by 256 B of mlps state table
    o = at("80 80 80 80 80 80 7b 7b 74 74 6f 6f")
    lps = u8(o, 4 * 64 * 2).reshape(4, 64, 2)
    assert (lps[:, :, 0] == lps[:, :, 1]).all()
    lps = lps[:, :, 0].T  # [state][q]
    assert lps[0].tolist() == [128, 176, 208, 240] and lps[62].tolist() == [6, 7, 8, 9]
    out["range_lps"] = lps
    mlps = u8(o + 512, 256)
    # index 128+s -> next s for MPS, 127-s -> next s for LPS, s = 2*state+mps
    next_mps = np.array([mlps[128 + 2 * s] >> 1 for s in range(64)])
    next_lps = np.array([mlps[127 - 2 * s] >> 1 for s in range(64)])
    assert next_mps[:4].tolist() == [1, 2, 3, 4] and next_mps[62] == 62
    assert next_lps[:8].tolist() == [0, 0, 1, 2, 2, 4, 4, 5]
    out["next_state_mps"] = next_mps
    out["next_state_lps"] = next_lps

    # --- CAVLC
    o_len = at("01 00 00 00 06 02 00 00 08 06 03 00 09 08 07 05")
    o_bits = at("01 00 00 00 05 01 00 00 07 04 01 00 07 06 05 03")
    out["coeff_token_len"] = u8(o_len, 4 * 68).reshape(4, 68)
    out["coeff_token_bits"] = u8(o_bits, 4 * 68).reshape(4, 68)
    assert out["coeff_token_len"][3][:8].tolist() == [6, 0, 0, 0, 6, 6, 0, 0]
    o_len = at("02 00 00 00 06 01 00 00 06 06 03 00")
    o_bits = at("01 00 00 00 07 01 00 00 04 06 01 00")
    out["chroma_dc_coeff_token_len"] = u8(o_len, 20)
    out["chroma_dc_coeff_token_bits"] = u8(o_bits, 20)
    o_len = at("01 03 03 04 04 05 05 06 06 07 07 08 08 09 09 09")
    o_bits = at("01 03 02 03 02 03 02 03 02 03 02 03 02 03 02 01")
    out["total_zeros_len"] = u8(o_len, 15 * 16).reshape(15, 16)
    out["total_zeros_bits"] = u8(o_bits, 15 * 16).reshape(15, 16)
    o_len = at("01 02 03 03 01 02 02 00 01 01 00 00")
    o_bits = at("01 01 01 00 01 01 00 00 01 00 00 00")
    out["chroma_dc_total_zeros_len"] = u8(o_len, 12).reshape(3, 4)
    out["chroma_dc_total_zeros_bits"] = u8(o_bits, 12).reshape(3, 4)
    o_len = at("01 01 00 00 00 00 00 00 00 00 00 00 00 00 00 00 01 02 02")
    o_bits = at("01 00 00 00 00 00 00 00 00 00 00 00 00 00 00 00 01 01 00 00 00 00 00 00 00 00 00 00 00 00 00 00 03 02 01 00")
    out["run_before_len"] = u8(o_len, 7 * 16).reshape(7, 16)
    out["run_before_bits"] = u8(o_bits, 7 * 16).reshape(7, 16)
    assert out["run_before_len"][6][:15].tolist() == [3, 3, 3, 3, 3, 3, 3, 4, 5, 6, 7, 8, 9, 10, 11]

    # --- CBP me(v): codeNum -> cbp; we need the inverse (cbp -> codeNum)
    intra = u8(at("2f 1f 0f 00 17 1b 1d 1e"), 48)
    inter = u8(at("00 10 01 02 04 08 20 03"), 48)
    inv_i = np.zeros(48, np.int32)
    inv_p = np.zeros(48, np.int32)
    inv_i[intra] = np.arange(48)
    inv_p[inter] = np.arange(48)
    assert sorted(intra.tolist()) == list(range(48)) and sorted(inter.tolist()) == list(range(48))
    out["cbp_to_codenum_intra"] = inv_i
    out["cbp_to_codenum_inter"] = inv_p

    # --- deblocking
    o = at("04 04 05 06 07 08 09 0a 0c 0d 0f 11 14 16 19 1c 20 24 28 2d 32 38 3f 47 50 5a 65 71 7f 90 a2 b6 cb e2 ff ff")
    alpha = np.concatenate([np.zeros(16, np.int32), u8(o, 36)])
    o = at("02 02 02 03 03 03 03 04 04 04 06 06 07 07 08 08 09 09 0a 0a 0b 0b 0c 0c 0d 0d 0e 0e 0f 0f 10 10 11 11 12 12")
    beta = np.concatenate([np.zeros(16, np.int32), u8(o, 36)])
    assert len(alpha) == 52 and len(beta) == 52 and alpha[51] == 255 and beta[51] == 18
    out["deblock_alpha"] = alpha
    out["deblock_beta"] = beta
    o = at("ff 00 00 01 ff 00 00 01 ff 00 00 01 ff 00 00 01 ff 00 01 01 ff 00 01 01 ff 01 01 01")
    tc = i8(o, 35 * 4).reshape(35, 4)[:, 1:]
    tc0 = np.concatenate([np.zeros((17, 3), np.int32), tc])
    assert tc0.shape == (52, 3) and tc0[51].tolist() == [13, 17, 25] and tc0[17].tolist() == [0, 0, 1]
    out["deblock_tc0"] = tc0

    # --- misc
    zz = u8(at("00 01 04 08 05 02 03 06 09 0c 0d 0a 07 0b 0e 0f"), 16)
    out["zigzag4x4"] = zz
    o = at("1c 1d 1d 1e 1f 20 20 21 22 22 23 23 24 24 25 25 25 26 26 26 27 27 27 27")
    cq = np.concatenate([np.arange(28), u8(o, 24)])
    assert len(cq) == 52 and cq[29] == 29 and cq[30] == 29 and cq[51] == 39
    out["chroma_qp"] = cq
    dq = u8(at("0a 0d 10 0b 0e 12 0d 10 14 0e 12 17 10 14 19 12 17 1d"), 18).reshape(6, 3)
    # stored (a, c, b): pos(0,0)-class, other-class, (1,1)-class  ->  emit as [a, b, c]
    out["dequant_v"] = dq[:, [0, 2, 1]]
    assert out["dequant_v"][0].tolist() == [10, 16, 13]
    # forward quant multipliers (encoder side; JM/x264 convention, MF*V*G = 2^21 approx)
    out["quant_mf"] = np.array([[13107, 5243, 8066], [11916, 4660, 7490], [10082, 4194, 6554],
                                [9362, 3647, 5825], [8192, 3355, 5243], [7282, 2893, 4559]])

    ctype = {"cabac_init_I": "int8_t", "cabac_init_P0": "int8_t", "deblock_tc0": "uint8_t",
             "quant_mf": "uint16_t"}

    def emit(name, a):
        a = np.asarray(a)
        t = ctype.get(name, "uint8_t")
        dims = "".join("[%d]" % s for s in a.shape)

        def fmt(x, ind):
            if x.ndim == 1:
                flat = x.tolist()
                if len(flat) <= 16:
                    return ind + "{" + ", ".join(str(v) for v in flat) + "}"
                rows = [ind + "    " + ", ".join(str(v) for v in flat[i:i + 16]) for i in range(0, len(flat), 16)]
                return ind + "{\n" + ",\n".join(rows) + "\n" + ind + "}"
            return ind + "{\n" + ",\n".join(fmt(y, ind + "    ") for y in x) + "\n" + ind + "}"

        return "H264_TABLE %s h264_%s%s =\n%s;\n" % (t, name, dims, fmt(a, ""))

    body = ["/* GENERATED by tools/extract_tables.py -- ITU-T H.264 standard tables. Do not edit. */",
            "#ifndef H264_TABLES_H", "#define H264_TABLES_H", "#include <stdint.h>",
            "#ifndef H264_TABLE", "#define H264_TABLE static const __attribute__((unused))", "#endif", ""]
    for k, v in out.items():
        body.append(emit(k, v))
    body.append("#endif /* H264_TABLES_H */\n")
    text = "\n".join(body)
    for rel in ("cedarx_h264_encoder_b200/csrc/h264_tables.h",):
        with open(os.path.join(ROOT, rel), "w") as f:
            f.write(text)
        print("wrote", rel, len(text), "bytes")


if __name__ == "__main__":
    main()
