"""GPU: the drop-in boundary. (1) the stand-alone CLI on the decoder.h-compatible shim; (2) the
reference's OWN unmodified main.cpp + parser.cpp linked against the shim in place of decoder.cpp /
cpuIDCT8x8.cpp / oclDCT8x8.cpp (oracle/_ref/ocljpegdec_b2j, built where the reference sources exist).
Both must write the BMP the reference writes: 54-byte header (decoder.cpp:372-395) + BGRA rows."""
import hashlib
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

import synth
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "ocljpegdecoder_b200", "bin", "b2jdec")
REF_MAIN = os.path.join(ROOT, "oracle", "_ref", "ocljpegdec_b2j")
REF_IDCT = os.path.join(ROOT, "oracle", "_ref", "ocljpegdec_b2jidct")


def _run(binary, jpeg_path, env_extra=None):
    with tempfile.TemporaryDirectory() as d:
        env = dict(os.environ)
        env.update(env_extra or {})
        out = subprocess.run([binary, jpeg_path], cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
        bmp = os.path.join(d, "m:\\output.bmp")
        data = open(bmp, "rb").read() if os.path.isfile(bmp) else None
        return out.stdout.decode(errors="replace"), data


def _check_bmp(bmp, width, height, want_bgra):
    assert bmp is not None and len(bmp) == 54 + width * height * 4
    magic, size, _, _, off = struct.unpack("<2sIHHI", bmp[:14])
    hsize, w, h, planes, bpp, comp = struct.unpack("<IiiHHI", bmp[14:34])
    assert (magic, size, off, hsize, w, h, planes, bpp, comp) == (b"BM", len(bmp), 54, 40, width, -height, 1, 32, 0)
    got = np.frombuffer(bmp[54:], np.uint8).reshape(height, width, 4)
    assert np.array_equal(got, want_bgra)


def test_standalone_cli_on_fixture(built, oracle):
    path = os.path.join(GOLDEN, "JPEG_example_JPG_RIP_050.jpg")
    log, bmp = _run(CLI, path)
    assert "decoding completed" in log, log
    rc, img, _, bgra = oracle.decode(open(path, "rb").read(), gate=0)
    _check_bmp(bmp, img.width, img.height, bgra)
    assert hashlib.sha256(bmp[54:]).hexdigest() == "efb49cf99f2f6c583c546d6ef24c5d26ae339955c8bbe467341933aafa9b16e5"


def test_standalone_cli_gate(built, oracle):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "x422.jpg")
        data = synth.synth_jpeg(130, 70, 3, 85, "422", 2)
        open(p, "wb").write(data)
        log, bmp = _run(CLI, p)                           # reference gate: 4:2:2 refused (decoder.cpp:58-69)
        assert bmp is None and "not supported" in log
        log, bmp = _run(CLI, p, {"B2J_GATE": "extended"})
        rc, img, _, bgra = oracle.decode(data, gate=1)
        _check_bmp(bmp, img.width, img.height, bgra)


@pytest.mark.skipif(not os.path.isfile(REF_MAIN), reason="oracle/_ref/ocljpegdec_b2j not built (needs /root/reference)")
def test_reference_main_and_parser_on_the_shim(built, oracle):
    for name, gate in (("JPEG_example_JPG_RIP_050", 0), ("g420_ri9_many", 0), ("g444_ri5_opt_q95", 0)):
        path = os.path.join(GOLDEN, name + ".jpg")
        log, bmp = _run(REF_MAIN, path)
        assert "decoding completed" in log and "End of Image" in log, log
        rc, img, _, bgra = oracle.decode(open(path, "rb").read(), gate=gate)
        _check_bmp(bmp, img.width, img.height, bgra)


@pytest.mark.skipif(not os.path.isfile(REF_IDCT), reason="oracle/_ref/ocljpegdec_b2jidct not built (needs /root/reference)")
def test_reference_decoder_on_the_clidct_shim(built, oracle):
    """The secondary boundary (idct.h:9-18): the reference's own main.cpp + parser.cpp + decoder.cpp, unmodified and built
    without USE_CPU_ONLY -- CPU Huffman, then clidct_create / allocate_memory / build / transfer / run / wait / retrieve /
    clean_up -- linked against csrc/refshim/idct_b2j.cpp in place of oclDCT8x8.cpp. The BMP must hold the pixels of the
    reference's CPU path."""
    for name in ("JPEG_example_JPG_RIP_050", "g444_ri5_opt_q95"):
        path = os.path.join(GOLDEN, name + ".jpg")
        log, bmp = _run(REF_IDCT, path)
        assert "clidct_run()" in log and "decoding completed" in log, log
        rc, img, _, bgra = oracle.decode(open(path, "rb").read(), gate=0)
        _check_bmp(bmp, img.width, img.height, bgra)
