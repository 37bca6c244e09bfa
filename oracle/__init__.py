"""oracle -- TEST INFRASTRUCTURE ONLY (ctypes bindings).

* ``Oracle``    : liboracle.so, the plain-C restatement of the reference CPU path (oracle.c).
* ``Reference`` : oracle/_ref/libjpegref.so, the UNMODIFIED reference CPU path compiled from
                  /root/reference/src by oracle/build_ref.sh (exists only where it was built or
                  where the prebuilt .so travelled to).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product (ocljpegdecoder_b200/) never does.
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

ORC_OK = 0
ORC_E_FORMAT = -1
ORC_E_UNSUPPORTED = -2
ORC_E_DATA = -3
GATE_REFERENCE = 0
GATE_EXTENDED = 1
GATE_GRAY = 4          # OR-able: one-component frames (extension beyond the reference, see oracle.c)


def build(with_ref=True):
    """Compile liboracle.so (and _ref/ when the reference sources are present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    if with_ref and os.path.isfile("/root/reference/src/decoder.cpp"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


class _OrcHuff(ctypes.Structure):
    _fields_ = [("num_codes", ctypes.c_int32), ("length", ctypes.c_uint8 * 256),
                ("code", ctypes.c_uint16 * 256), ("value", ctypes.c_uint8 * 256)]


class OrcImage(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32), ("height", ctypes.c_int32),
        ("sampling", ctypes.c_int32 * 3), ("quant_id", ctypes.c_int32 * 3), ("huff_id", ctypes.c_int32 * 3),
        ("restart_interval", ctypes.c_int32),
        ("mcu_width", ctypes.c_int32), ("mcu_height", ctypes.c_int32),
        ("mcu_count_w", ctypes.c_int32), ("mcu_count_h", ctypes.c_int32), ("mcu_count", ctypes.c_int32),
        ("blks_per_mcu", ctypes.c_int32 * 3), ("tot_blks_per_mcu", ctypes.c_int32), ("blk_count", ctypes.c_int32),
        ("scan_offset", ctypes.c_int64),
        ("quant_present", ctypes.c_int32 * 4), ("quant", (ctypes.c_int32 * 64) * 4),
        ("huff_present", ctypes.c_int32 * 32), ("huff", _OrcHuff * 32),
    ]


class Oracle:
    """The C restatement. decode() returns (rc, info, coef int32[blk,64], bgra uint8[H,W,4])."""

    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.isfile(path):
            build(with_ref=False)
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.orc_parse.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(OrcImage)]
        L.orc_huffman.argtypes = [ctypes.POINTER(OrcImage), ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]
        L.orc_pixels.argtypes = [ctypes.POINTER(OrcImage), ctypes.c_void_p, ctypes.c_void_p]
        L.orc_idct.argtypes = [ctypes.c_void_p]
        L.orc_yuv_to_rgb32.argtypes = [ctypes.c_int32] * 3
        L.orc_yuv_to_rgb32.restype = ctypes.c_uint32
        L.orc_zigzag.restype = ctypes.POINTER(ctypes.c_int * 64)
        L.orc_sizeof_image.restype = ctypes.c_size_t
        assert L.orc_sizeof_image() == ctypes.sizeof(OrcImage)

    def set_strict(self, strict):
        """strict=True (default): reproduce the reference's lost-RSTn-at-chunk-end defect
        (decoder.cpp:118-131); strict=False: the intended behaviour (what the GPU path implements)."""
        self.lib.orc_set_strict(1 if strict else 0)

    def zigzag(self):
        return np.array(list(self.lib.orc_zigzag().contents), dtype=np.int32)

    def parse(self, data, gate=GATE_EXTENDED):
        img = OrcImage()
        rc = self.lib.orc_parse(data, len(data), gate, ctypes.byref(img))
        return rc, img

    def decode(self, data, gate=GATE_EXTENDED, want_pixels=True):
        rc, img = self.parse(data, gate)
        if rc != ORC_OK:
            return rc, img, None, None
        coef = np.zeros((img.blk_count, 64), dtype=np.int32)
        rc = self.lib.orc_huffman(ctypes.byref(img), data, len(data), coef.ctypes.data)
        if rc != ORC_OK:
            return rc, img, None, None
        bgra = None
        if want_pixels:
            work = coef.copy()
            bgra = np.zeros((img.height, img.width, 4), dtype=np.uint8)
            rc = self.lib.orc_pixels(ctypes.byref(img), work.ctypes.data, bgra.ctypes.data)
        return rc, img, coef, bgra

    def idct(self, block):
        b = np.ascontiguousarray(block, dtype=np.int32).reshape(64).copy()
        self.lib.orc_idct(b.ctypes.data)
        return b

    def yuv_to_rgb32(self, y, u, v):
        return self.lib.orc_yuv_to_rgb32(int(y), int(u), int(v))


class RefInfo(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32), ("height", ctypes.c_int32),
        ("mcu_width", ctypes.c_int32), ("mcu_height", ctypes.c_int32),
        ("mcu_count_w", ctypes.c_int32), ("mcu_count_h", ctypes.c_int32),
        ("blks_per_mcu", ctypes.c_int32 * 3), ("tot_blks_per_mcu", ctypes.c_int32),
        ("blk_count", ctypes.c_int32), ("restart_interval", ctypes.c_int32),
        ("sampling", ctypes.c_int32 * 3), ("scan_offset", ctypes.c_int64),
    ]


def reference_available():
    return os.path.isfile(os.path.join(_HERE, "_ref", "libjpegref.so"))


class Reference:
    """The unmodified reference CPU path (Huffman + cpuIDCT8x8 + YUV_to_RGB32 + BMP)."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libjpegref.so")
        if not os.path.isfile(path):
            raise FileNotFoundError(path + " (run oracle/build_ref.sh where /root/reference exists)")
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.ref_open.restype = ctypes.c_void_p
        L.ref_open.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(RefInfo)]
        L.ref_huffman.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
        L.ref_pixels.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
        L.ref_close.argtypes = [ctypes.c_void_p]
        L.ref_fast_idct.argtypes = [ctypes.c_void_p]
        L.ref_yuv_to_rgb32.argtypes = [ctypes.c_int32] * 3
        L.ref_yuv_to_rgb32.restype = ctypes.c_uint32
        self._tmp = None

    def workdir(self):
        if self._tmp is None:
            # the reference writes an uncompressed BMP per image: prefer tmpfs when it has room
            base = None
            try:
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize > (2 << 30):
                    base = "/dev/shm"
            except OSError:
                pass
            self._tmp = tempfile.TemporaryDirectory(prefix="jpegref_", dir=base)
        return self._tmp.name

    def decode(self, data, skip_gate=False, want_pixels=True):
        """Returns (ok, info, coef, bgra, (t_huffman_s, t_mcu_s)). ok False <=> the reference refused/failed."""
        info = RefInfo()
        h = self.lib.ref_open(data, len(data), 1 if skip_gate else 0, ctypes.byref(info))
        if not h:
            return False, None, None, None, (0.0, 0.0)
        try:
            coef = np.zeros((info.blk_count, 64), dtype=np.int32)
            t1 = ctypes.c_double()
            t2 = ctypes.c_double()
            if not self.lib.ref_huffman(h, coef.ctypes.data, ctypes.byref(t1)):
                return False, info, None, None, (t1.value, 0.0)
            bgra = None
            if want_pixels:
                bgra = np.zeros((info.height, info.width, 4), dtype=np.uint8)
                if not self.lib.ref_pixels(h, self.workdir().encode(), bgra.ctypes.data, ctypes.byref(t2)):
                    return False, info, coef, None, (t1.value, t2.value)
            return True, info, coef, bgra, (t1.value, t2.value)
        finally:
            self.lib.ref_close(h)

    def fast_idct(self, block):
        b = np.ascontiguousarray(block, dtype=np.int32).reshape(64).copy()
        self.lib.ref_fast_idct(b.ctypes.data)
        return b

    def yuv_to_rgb32(self, y, u, v):
        return self.lib.ref_yuv_to_rgb32(int(y), int(u), int(v))
