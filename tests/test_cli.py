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
    for flags in (["--bogus"], ["--gpus", "0"], ["--gpus", "65"], ["--batch-gops", "2", "--queue-gops", "1"], ["--handles", "-1"]):
        r = subprocess.run([CLI, "-", "64", "48", out] + flags, input=b"", capture_output=True)
        assert r.returncode != 0 and r.stderr, flags


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
