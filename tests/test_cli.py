"""The CLI surface: `h264enc <infile or -> <width> <height> <outfile>` (userspace/h264enc.c:141-147)."""
import os
import subprocess

import pytest
import torch

from common import make_clip, oracle_encode_clip

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cedarx_h264_encoder_b200", "h264enc")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "h264enc")


def test_usage_matches_reference(product_lib):
    """argc != 5: 'Usage: %s <infile> <width> <height> <outfile>' and exit status 255 (return -1)."""
    r = subprocess.run([CLI, "a", "b"], capture_output=True, text=True)
    assert r.returncode == 255
    assert r.stdout.strip() == "Usage: %s <infile> <width> <height> <outfile>" % CLI
    if os.path.exists(REF_CLI):  # the reference's own program, compiled from /root/reference by oracle/Makefile
        q = subprocess.run([REF_CLI, "a", "b"], capture_output=True, text=True)
        assert q.returncode == r.returncode
        assert q.stdout.replace(REF_CLI, "X") == r.stdout.replace(CLI, "X")


def test_cli_rejects_bad_extension_flags(product_lib, tmp_path):
    """unknown flags and impossible combinations end with a message and a non-zero status before anything is opened"""
    out = str(tmp_path / "o.264")
    for flags in (["--bogus"], ["--gpus", "0"], ["--gpus", "65"], ["--batch-gops", "2", "--queue-gops", "1"], ["--handles", "-1"],
                  ["--reader-threads", "0"], ["--reader-threads", "17"]):
        r = subprocess.run([CLI, "-", "64", "48", out] + flags, input=b"", capture_output=True)
        assert r.returncode != 0 and r.stderr, flags


def test_parallel_batch_reader(tmp_path):
    """--reader-threads: a batch of a regular file is read with pread() by several threads (h264enc.c
    read_batch_parallel).  The same function, compiled into a self-test program: every batch size and thread count
    returns the file's whole frames in order, the trailing partial frame is dropped (userspace/h264enc.c:181-187 stops
    at the first short read), and the file position ends behind the last whole frame."""
    exe = str(tmp_path / "reader_selftest")
    subprocess.run(["gcc", "-O2", "-std=gnu11", "-DH264ENC_READER_SELFTEST", "-I", os.path.join(ROOT, "include"), "-pthread",
                    "-o", exe, os.path.join(ROOT, "cedarx_h264_encoder_b200", "csrc", "h264enc.c")], check=True)
    fb, nfr = 4099, 37
    data = os.urandom(fb * nfr + 1234)
    f = tmp_path / "in.bin"
    f.write_bytes(data)
    for threads in (1, 2, 3, 5, 16):
        for cap in (1, 7, 37, 64):
            r = subprocess.run([exe, str(f), str(fb), str(cap), str(threads)], capture_output=True)
            assert r.returncode == 0
            assert r.stdout == data[:fb * nfr], (threads, cap)
            counts = [int(x) for x in r.stderr.split()]
            assert sum(counts) == nfr and all(c == cap for c in counts[:-1]) and counts[-1] < cap
    for content in (b"", b"abc"):
        f.write_bytes(content)
        r = subprocess.run([exe, str(f), "10", "4", "3"], capture_output=True)
        assert r.returncode == 0 and r.stdout == b""


def test_cli_fails_loudly_without_gpu(product_lib, tmp_path):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([CLI, "-", "64", "48", str(tmp_path / "o.264")], input=b"", capture_output=True)
    assert r.returncode != 0
    assert b"no usable CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available(), reason="no CUDA device")
@pytest.mark.parametrize("flags,cfg", [([], dict(qp=24, gop=25, cabac=1)),
                                       (["--cavlc", "--qp", "30", "--gop", "4"], dict(qp=30, gop=4, cabac=0)),
                                       (["--gop", "2", "--batch-gops", "2", "--slice-rows", "2"], dict(qp=24, gop=2, cabac=1, slice_rows=2)),
                                       (["--gop", "2", "--batch-gops", "1", "--handles", "3"], dict(qp=24, gop=2, cabac=1)),
                                       (["--gop", "1", "--batch-gops", "1", "--handles", "1"], dict(qp=24, gop=1, cabac=1)),
                                       (["--gop", "2", "--queue-gops", "1"], dict(qp=24, gop=2, cabac=1)),
                                       (["--gop", "3", "--queue-gops", "4", "--cavlc"], dict(qp=24, gop=3, cabac=0)),
                                       (["--gop", "2", "--gpus", "1", "--batch-gops", "1"], dict(qp=24, gop=2, cabac=1)),
                                       (["--gop", "2", "--devices", "0,0"], dict(qp=24, gop=2, cabac=1))])
def test_cli_pipe_equals_golden_model(product_lib, tmp_path, flags, cfg):
    """raw nv12 on stdin -> Annex-B file, reference defaults; trailing garbage (short frame) is ignored
    exactly like the reference's read loop (userspace/h264enc.c:183-187)."""
    w, h, n = 86, 50, 6
    clip = make_clip("synth", w, h, n)
    want, sizes, _ = oracle_encode_clip(clip, w, h, me_range=16, **cfg)
    out = tmp_path / "o.264"
    r = subprocess.run([CLI, "-", str(w), str(h), str(out)] + flags, input=clip.tobytes() + b"\x00" * 100,
                       capture_output=True)
    assert r.returncode == 0, r.stderr
    assert out.read_bytes() == want
    assert ("Frame %5d: %5dbytes" % (n - 1, sizes[-1])).encode() in r.stdout
    assert oct(out.stat().st_mode & 0o777) == oct(0o644)


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available(), reason="no CUDA device")
@pytest.mark.parametrize("flags", [[], ["--reader-threads", "1"], ["--reader-threads", "3"], ["--handles", "3", "--reader-threads", "16"]])
def test_cli_file_input_with_parallel_reader_equals_golden_model(product_lib, tmp_path, flags):
    """the pipelined modes read a regular input file batch by batch with several threads; same bytes, same progress
    lines, and the partial frame at the end of the file is ignored"""
    w, h, n = 86, 50, 11
    clip = make_clip("synth", w, h, n)
    want, sizes, _ = oracle_encode_clip(clip, w, h, me_range=16, qp=24, gop=2, cabac=1)
    src, out = tmp_path / "in.yuv", tmp_path / "o.264"
    src.write_bytes(clip.tobytes() + b"\x00" * 1000)
    r = subprocess.run([CLI, str(src), str(w), str(h), str(out), "--gop", "2", "--batch-gops", "2"] + flags, capture_output=True)
    assert r.returncode == 0, r.stderr
    assert out.read_bytes() == want
    assert ("Frame %5d: %5dbytes" % (n - 1, sizes[-1])).encode() in r.stdout
