"""The product's ordered multi-handle / multi-GPU pipeline (csrc/pipeline.cpp: cedar_b200_pipe_*, SURVEY 8e "host thread
per GPU, per-GPU bytestreams concatenated on the host") without a GPU: pipeline.cpp only uses the public C ABI, so
tests/pipeline_harness.cpp links it against a stand-in for the seven encoder calls it makes and drives it from a producer
and a consumer thread.  Checked: batches come back in submission order with their position in the stream
(first_frame_index: IDR at every multiple of the keyframe interval and parameter sets once, kernel/cedar.c:1047-1061,
1193-1196), only the last batch may be short, both ways of ending a stream, polling and blocking consumers, a failing
batch is reported in its place and nothing else, open failures close every handle again, close with batches still
queued, and the argument errors -- under ThreadSanitizer and under AddressSanitizer + UBSan.  (The same calls on the
real encoder: tests/test_gpu_parity.py::test_pipeline_*.)"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("san", ["-fsanitize=thread", "-fsanitize=address,undefined -fno-sanitize-recover=undefined"])
def test_pipeline_state_machine_under_sanitizers(tmp_path, san):
    exe = str(tmp_path / "pipeline_harness")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-Wall", "-Wextra", "-pthread", *san.split(), "-o", exe,
                    os.path.join(ROOT, "tests", "pipeline_harness.cpp"),
                    os.path.join(ROOT, "cedarx_h264_encoder_b200", "csrc", "pipeline.cpp")], check=True)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1", ASAN_OPTIONS="detect_leaks=1", UBSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([exe, "4"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == "pipeline_harness ok", r.stderr[-3000:]
