"""The drop-in boundary: the C-ABI library loads and exports every symbol include/cedar_b200.h
declares; config validation follows kernel/cedar.c:744-789; without a GPU open() fails loudly with
-ENODEV (there is no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cedar_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cedar_b200_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_three_reference_calls():
    syms = declared_symbols()
    for s in ("cedar_b200_open", "cedar_b200_encode_frame", "cedar_b200_close"):
        assert s in syms


def test_library_exports_every_declared_symbol(product_lib):
    for s in declared_symbols():
        assert hasattr(product_lib, s), "missing export: " + s


def test_config_struct_mirrors_cedar_ioctl_config(product_lib):
    from cedarx_h264_encoder_b200 import api
    names = [f[0] for f in api.CedarConfig._fields_]
    # kernel/cedar_ioctl.h:12-32, in order
    assert names[:12] == ["src_width", "src_height", "src_format", "dst_width", "dst_height", "profile", "level", "qp",
                          "keyframe_interval", "thumbnail", "thumbnail_downscale", "entropy_coding_mode"]
    ionames = [f[0] for f in api.CedarIO._fields_]
    assert ionames == ["input_luma", "input_luma_size", "input_chroma", "input_chroma_size", "bytestream",
                       "bytestream_size"]


def _open_rc(product_lib, **kw):
    from cedarx_h264_encoder_b200 import api
    base = dict(width=64, height=48)
    base.update(kw)
    cfg = api.make_config(**base)
    io, h = api.CedarIO(), C.c_void_p()
    r = product_lib.cedar_b200_open(C.byref(cfg), C.byref(io), C.byref(h))
    if r == 0:
        product_lib.cedar_b200_close(h)
    return r


def test_validation_matches_reference_rules(product_lib):
    """Validation runs before the device is touched, so the -EINVAL cases are checkable anywhere."""
    EINVAL = -22
    assert _open_rc(product_lib, width=63) == EINVAL
    assert _open_rc(product_lib, height=47) == EINVAL
    assert _open_rc(product_lib, dst_width=70) == EINVAL
    assert _open_rc(product_lib, width=80, dst_width=64) == EINVAL
    assert _open_rc(product_lib, qp=0) == EINVAL and _open_rc(product_lib, qp=48) == EINVAL
    assert _open_rc(product_lib, fmt=2) == EINVAL
    assert _open_rc(product_lib, gop=0) == EINVAL
    assert _open_rc(product_lib, gop=32, relax_gop=0) == EINVAL
    assert _open_rc(product_lib, me_range=65) == EINVAL


def test_no_cpu_fallback(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _open_rc(product_lib) == -19  # -ENODEV, and a message on stderr


def test_encode_frame_rejects_null_handle(product_lib):
    assert product_lib.cedar_b200_encode_frame(None) == -22  # cedar.c:1039-1043: not configured -> -EINVAL


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from cedarx_h264_encoder_b200 import api
    monkeypatch.setattr(api, "_lib", None)
    monkeypatch.setattr(api, "PKG_DIR", str(tmp_path))
    with pytest.raises(api.LibraryMissing):
        api.load_library()


def test_pipeline_fails_loudly_without_gpu_and_rejects_bad_arguments(product_lib):
    """cedar_b200_pipe_open: argument errors are -EINVAL anywhere; on a box without a CUDA device every worker's open fails
    with -ENODEV and the call returns it (no hang, nothing left behind)."""
    import torch
    from cedarx_h264_encoder_b200 import api
    cfg = api.make_config(64, 48)
    p = C.c_void_p()
    assert product_lib.cedar_b200_pipe_open(C.byref(cfg), None, -1, 0, 0, C.byref(p)) == -22
    assert product_lib.cedar_b200_pipe_open(None, None, 0, 0, 0, C.byref(p)) == -22
    bad = api.make_config(64, 48, qp=0)
    assert product_lib.cedar_b200_pipe_open(C.byref(bad), None, 0, 2, 1, C.byref(p)) == -22
    if not torch.cuda.is_available():
        assert product_lib.cedar_b200_pipe_open(C.byref(cfg), None, 0, 2, 1, C.byref(p)) == -19
    assert product_lib.cedar_b200_flush(None) == -22
    assert product_lib.cedar_b200_pipe_next(None, None, None, None, None, 0) == -22
