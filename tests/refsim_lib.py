"""ctypes wrapper of oracle/_ref/librefsim.so: the reference's kernel driver (kernel/cedar.c, compiled unmodified by
oracle/Makefile from /root/reference) running in user space against a software model of the video engine
(oracle/refsim/).  TEST INFRASTRUCTURE ONLY.

The library is built in the container that has /root/reference and travels to the GPU box prebuilt (oracle/_ref/ is
git-ignored, not gpurun-ignored); nothing here reads /root/reference at run time."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_ref", "librefsim.so")
CLI = os.path.join(ORACLE_DIR, "_ref", "h264enc_sim")
CLI_B200 = os.path.join(ORACLE_DIR, "_ref", "h264enc_b200")  # the same program bound to the product (refsim/b200_shim.c)
REF = os.environ.get("CEDAR_REFERENCE", "/root/reference")

IOCTL_ENCODE, IOCTL_CONFIG = 0x600, 0x601  # enum cedar_ioctl_cmd, kernel/cedar_ioctl.h:7-10
ENC_BASE, ISP_BASE = 0xB00, 0xA00          # kernel/cedar_regs.h:7,31
PARA0, PARA1, MEPARA = 0x04, 0x08, 0x10    # kernel/cedar_regs.h:34-37


class IoctlConfig(C.Structure):
    """struct cedar_ioctl_config, kernel/cedar_ioctl.h:12-46."""
    _fields_ = [(n, C.c_int) for n in ("src_width", "src_height", "src_format", "dst_width", "dst_height", "profile", "level",
                                        "qp", "keyframe_interval", "thumbnail", "thumbnail_downscale", "entropy_coding_mode")] + \
               [("input_luma_dma_addr", C.c_uint32), ("input_luma_size", C.c_int),
                ("input_chroma_dma_addr", C.c_uint32), ("input_chroma_size", C.c_int),
                ("bytestream_dma_addr", C.c_uint32), ("bytestream_size", C.c_int),
                ("thumb_luma_dma_addr", C.c_uint32), ("thumb_luma_size", C.c_int),
                ("thumb_chroma_dma_addr", C.c_uint32), ("thumb_chroma_size", C.c_int)]


_lib = None


def available():
    return os.path.exists(LIB) or os.path.exists(os.path.join(REF, "kernel", "cedar.c"))


def lib():
    global _lib
    if _lib is None:
        if os.path.exists(os.path.join(REF, "kernel", "cedar.c")):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "refsim", "REF=" + REF])
        L = C.CDLL(LIB)
        L.refsim_ioctl.argtypes = [C.c_uint, C.c_void_p]
        L.refsim_ioctl.restype = C.c_long
        L.refsim_mmap.argtypes = [C.c_size_t, C.c_uint32]
        L.refsim_mmap.restype = C.c_void_p
        L.refsim_state.argtypes = [C.c_char_p]
        L.refsim_state.restype = C.c_long
        L.refsim_reg.argtypes = [C.c_int]
        L.refsim_reg.restype = C.c_uint32
        L.refsim_log.restype = C.c_char_p
        L.refsim_ve_fault.restype = C.c_char_p
        _lib = L
    return _lib


def make_config(width, height, qp=24, gop=25, cabac=1, fmt=0, profile=77, level=41, dst_width=None, dst_height=None,
                thumbnail=0, thumbnail_downscale=0):
    """Defaults: what userspace/h264enc.c:53-66 hard-codes."""
    c = IoctlConfig()
    c.src_width, c.src_height, c.src_format = width, height, fmt
    c.dst_width = (width + 15) & ~15 if dst_width is None else dst_width
    c.dst_height = (height + 15) & ~15 if dst_height is None else dst_height
    c.profile, c.level, c.qp, c.keyframe_interval = profile, level, qp, gop
    c.thumbnail, c.thumbnail_downscale, c.entropy_coding_mode = thumbnail, thumbnail_downscale, cabac
    return c


class Device:
    """One open of /dev/cedar_dev in the simulation (the driver allows one opener at a time)."""

    def __init__(self, me_range=16):
        self.L = lib()
        self.L.refsim_set_me_range(me_range)
        r = self.L.refsim_open()
        if r:
            raise OSError(-r, "refsim_open: %s" % os.strerror(-r))
        self.opened = True
        self.cfg = None

    def config(self, cfg: IoctlConfig):
        """ioctl(CEDAR_IOCTL_CONFIG) + the three mmaps of userspace/h264enc.c:68-106.  Returns the ioctl's value."""
        r = self.L.refsim_ioctl(IOCTL_CONFIG, C.addressof(cfg))
        if r == 0:
            self.cfg = cfg
            self.luma = self._map(cfg.input_luma_size, cfg.input_luma_dma_addr)
            self.chroma = self._map(cfg.input_chroma_size, cfg.input_chroma_dma_addr)
            self.bytestream = self._map(cfg.bytestream_size, cfg.bytestream_dma_addr)
        return int(r)

    def _map(self, size, addr):
        p = self.L.refsim_mmap(size, addr)
        if not p:
            raise OSError("refsim_mmap(%d, 0x%08x) failed: %s" % (size, addr, self.log()))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(size,))

    def encode(self, luma, chroma) -> bytes:
        """read_frame x2 + ioctl(CEDAR_IOCTL_ENCODE) + write(ret bytes), userspace/h264enc.c:178-198."""
        y = np.asarray(luma, np.uint8).reshape(-1)
        c = np.asarray(chroma, np.uint8).reshape(-1)
        self.luma[:y.size] = y
        self.chroma[:c.size] = c
        r = self.L.refsim_ioctl(IOCTL_ENCODE, None)
        if r < 0:
            raise OSError("CEDAR_IOCTL_ENCODE returned %d: %s / %s" % (r, self.L.refsim_ve_fault().decode(), self.log()))
        return self.bytestream[:r].tobytes()

    def state(self, name):
        return int(self.L.refsim_state(name.encode()))

    def reg(self, offset):
        return int(self.L.refsim_reg(offset))

    def log(self):
        return self.L.refsim_log().decode(errors="replace")

    def close(self):
        if self.opened:
            self.luma = self.chroma = self.bytestream = None
            self.L.refsim_release()
            self.opened = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def split_nals(stream: bytes):
    """Annex-B with the reference's 4-byte start codes (cedar.c:872-880) -> list of NAL units (header byte first)."""
    out, i = [], 0
    marks = []
    while True:
        j = stream.find(b"\x00\x00\x00\x01", i)
        if j < 0:
            break
        marks.append(j)
        i = j + 4
    for k, j in enumerate(marks):
        out.append(stream[j + 4:marks[k + 1] if k + 1 < len(marks) else len(stream)])
    return out


def rbsp_bits(nal: bytes, nbits: int) -> str:
    """First nbits of the NAL's payload (after the header byte), emulation prevention removed."""
    out, zeros = [], 0
    for b in nal[1:]:
        if zeros >= 2 and b == 3:
            zeros = 0
            continue
        out.append(b)
        zeros = zeros + 1 if b == 0 else 0
        if len(out) * 8 >= nbits:
            break
    return "".join(format(b, "08b") for b in out)[:nbits]
