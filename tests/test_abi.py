"""CPU: the C-ABI library loads, exports every symbol include/b2j.h declares, parses headers like
the oracle does, and refuses to run without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import synth
from conftest import ROOT


def test_library_exports_every_declared_symbol(built):
    import ocljpegdecoder_b200 as b2j
    with open(os.path.join(ROOT, "include", "b2j.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(b2j_[a-z_0-9]+)\s*\(", src)))
    assert declared, "no declarations found"
    L = ctypes.CDLL(b2j.library_path())
    for name in declared:
        assert hasattr(L, name), "missing export " + name
    assert sorted(declared) == sorted(b2j.EXPORTED_SYMBOLS)
    assert L.b2j_abi_version() == 5


def test_struct_sizes_match_header(built):
    import subprocess
    import tempfile
    import ocljpegdecoder_b200 as b2j
    code = '#include "b2j.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu", sizeof(b2j_image_desc), sizeof(b2j_batch_info), sizeof(b2j_stage_times), sizeof(b2j_host_opts));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        with open(c, "w") as f:
            f.write(code)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "s")])
        out = subprocess.check_output([os.path.join(d, "s")]).decode().split()
    assert [int(x) for x in out] == [ctypes.sizeof(b2j.ImageDesc), ctypes.sizeof(b2j.BatchInfo), ctypes.sizeof(b2j.StageTimes), ctypes.sizeof(b2j.HostOpts)]


def test_parse_header_agrees_with_oracle(built, oracle, fixture_jpeg):
    import ocljpegdecoder_b200 as b2j
    files = [fixture_jpeg]
    for i, (w, h, ss, q, ri, opt) in enumerate([(64, 48, "444", 75, 0, False), (67, 45, "420", 90, 7, True), (130, 70, "422", 85, 2, False)]):
        files.append(synth.synth_jpeg(w, h, 40 + i, q, ss, ri, opt))
    for data in files:
        for gate in (0, 1):
            rc_o, img = oracle.parse(data, gate)
            rc, d = b2j.parse_header(data, gate)
            assert (rc == 0) == (rc_o == 0)
            if rc != 0:
                continue
            assert (d.width, d.height, d.restart_interval) == (img.width, img.height, img.restart_interval)
            assert (d.mcu_width, d.mcu_height, d.mcu_count_w, d.mcu_count_h, d.mcu_count) == (
                img.mcu_width, img.mcu_height, img.mcu_count_w, img.mcu_count_h, img.mcu_count)
            assert list(d.blks_per_mcu) == list(img.blks_per_mcu) and d.blk_count == img.blk_count
            assert d.scan_offset == img.scan_offset and d.scan_offset + d.scan_size == len(data)
            for c in range(3):
                assert d.sampling[c] == img.sampling[c] and d.quant_id[c] == img.quant_id[c] and d.huff_id[c] == img.huff_id[c]
                assert list(d.quant[d.quant_id[c]]) == list(img.quant[img.quant_id[c]])


def test_parse_header_rejects_like_reference(built, oracle, fixture_jpeg):
    import ocljpegdecoder_b200 as b2j
    bad = [b"", b"\xff\xd8", fixture_jpeg[:200], b"\x00" + fixture_jpeg[1:], fixture_jpeg.replace(b"\xff\xc0", b"\xff\xc2", 1)]
    # a COM segment before the tables: load_jpg() stops at unknown markers (parser.cpp:410-412)
    bad.append(fixture_jpeg[:2] + b"\xff\xfe\x00\x04ab" + fixture_jpeg[2:])
    for data in bad:
        rc_o, _ = oracle.parse(data, 1)
        rc, _ = b2j.parse_header(data, 1)
        assert rc != 0 and rc_o != 0


def test_no_device_no_fallback(built):
    import ocljpegdecoder_b200 as b2j
    L = b2j.load_library()
    if L.b2j_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(b2j.B2JError) as ei:
        b2j.Decoder(0)
    assert ei.value.code == -7   # B2J_E_NODEVICE: the product path fails loudly without the device


def test_robust_parse_mode_accepts_what_the_reference_parser_chokes_on(built, oracle, fixture_jpeg):
    """COM / late APPn segments end the reference's parsing (parser.cpp:410-412).
    With B2J_PARSE_ROBUST they are skipped and the header equals that of the clean file."""
    import ocljpegdecoder_b200 as b2j
    base = synth.synth_jpeg(96, 64, 77, 85, "420", 4)
    rc0, d0 = b2j.parse_header(base, b2j.GATE_REFERENCE)
    assert rc0 == 0
    com = b"\xff\xfe" + (2 + 11).to_bytes(2, "big") + b"hello world"
    app1 = b"\xff\xe1" + (2 + 6).to_bytes(2, "big") + b"Exif\x00\x00"
    k = base.index(b"\xff\xdb")                      # in front of the first DQT
    k2 = base.index(b"\xff\xc4")                     # in front of the first DHT
    dirty = base[:k] + com + base[k:k2] + app1 + b"\xff" + base[k2:]     # + one fill byte before a marker
    assert b2j.parse_header(dirty, b2j.GATE_REFERENCE)[0] != 0 and oracle.parse(dirty, 0)[0] != 0
    rc, d = b2j.parse_header(dirty, b2j.GATE_REFERENCE | b2j.PARSE_ROBUST)
    assert rc == 0
    extra = len(dirty) - len(base)
    assert (d.width, d.height, d.blk_count, d.restart_interval) == (d0.width, d0.height, d0.blk_count, d0.restart_interval)
    assert d.scan_offset == d0.scan_offset + extra and d.scan_size == d0.scan_size
    assert bytes(d.huff_counts) == bytes(d0.huff_counts) and bytes(d.quant) == bytes(d0.quant)
    # robust mode is a superset: clean files parse identically
    rc1, d1 = b2j.parse_header(base, b2j.GATE_EXTENDED | b2j.PARSE_ROBUST)
    assert rc1 == 0 and bytes(d1.huff_symbols) == bytes(d0.huff_symbols) and d1.scan_offset == d0.scan_offset
    assert b2j.parse_header(fixture_jpeg, b2j.PARSE_ROBUST)[0] == 0
    # progressive stays rejected
    assert b2j.parse_header(base.replace(b"\xff\xc0", b"\xff\xc2", 1), b2j.PARSE_ROBUST)[0] != 0


def test_grayscale_gate_flag(built, oracle):
    """One-component frames are refused like the reference refuses them unless B2J_GATE_GRAY is set; with it the
    geometry is one 8x8 block per MCU and agrees with the oracle's extension."""
    import io
    from PIL import Image
    import ocljpegdecoder_b200 as b2j
    from oracle import GATE_GRAY, GATE_EXTENDED
    buf = io.BytesIO()
    Image.fromarray(synth.synth_pixels(77, 45, 1)[:, :, 0], "L").save(buf, format="JPEG", quality=80)
    f = buf.getvalue()
    for gate in (b2j.GATE_REFERENCE, b2j.GATE_EXTENDED, b2j.GATE_EXTENDED | b2j.PARSE_ROBUST):
        assert b2j.parse_header(f, gate)[0] != 0
    rc, d = b2j.parse_header(f, b2j.GATE_EXTENDED | b2j.GATE_GRAY)
    assert rc == 0
    rco, o = oracle.parse(f, GATE_EXTENDED | GATE_GRAY)
    assert rco == 0
    assert (d.width, d.height, d.mcu_width, d.mcu_height, d.mcu_count, d.tot_blks_per_mcu, d.blk_count) == \
           (o.width, o.height, o.mcu_width, o.mcu_height, o.mcu_count, o.tot_blks_per_mcu, o.blk_count) == (77, 45, 8, 8, 60, 1, 60)
    assert d.scan_offset == o.scan_offset
    # a colour file is unaffected by the flag
    c = synth.synth_jpeg(64, 48, 2, 90, "420", 0)
    assert b2j.parse_header(c, b2j.GATE_REFERENCE | b2j.GATE_GRAY)[0] == 0


def test_file_cut_behind_the_sos_header_is_refused(built, fixture_jpeg):
    """A file that ends right behind the SOS header has no entropy-coded data: the reference gives up on it with
    "data incomplete" (decoder.cpp:310-314); here the header parser reports B2J_E_DATA and no batch is built."""
    import ocljpegdecoder_b200 as b2j
    rc, d = b2j.parse_header(fixture_jpeg, b2j.GATE_REFERENCE)
    assert rc == 0 and d.scan_size > 0
    cut = fixture_jpeg[:d.scan_offset]
    rc, d2 = b2j.parse_header(cut, b2j.GATE_REFERENCE)
    assert rc == -4                      # B2J_E_DATA
    rc, _ = b2j.parse_header(fixture_jpeg[:d.scan_offset + 1], b2j.GATE_REFERENCE)
    assert rc == 0                       # one byte of scan: a batch can be built, the decode flags it


def test_huffman_table_ids_above_3_are_refused(built, oracle):
    """Documented deviation: the reference keeps 16 DC + 16 AC table slots (parser.cpp:176), the C ABI
    carries ids 0..3 (all that baseline JPEG allows, ITU T.81 B.2.4.2). A DHT with id 4 is a format error here."""
    import jpegcraft
    import ocljpegdecoder_b200 as b2j
    blocks = np.zeros((6, 64), np.int64)
    tables = dict(jpegcraft.STD_TABLES)
    tables[(0, 4)] = tables[(0, 0)]
    data = jpegcraft.build_jpeg(16, 16, (2, 2), blocks, [[1] * 64, [1] * 64], tables=tables)
    rc, _ = b2j.parse_header(data, b2j.GATE_REFERENCE)
    assert rc == -2                      # B2J_E_FORMAT
    # a scan that SELECTS table 4 is refused as well, whatever DHTs are present
    good = jpegcraft.build_jpeg(16, 16, (2, 2), blocks, [[1] * 64, [1] * 64])
    k = good.index(b"\xff\xda")
    bad = bytearray(good)
    bad[k + 6] = 0x40                    # component 1: Td = 4, Ta = 0
    rc, _ = b2j.parse_header(bytes(bad), b2j.GATE_REFERENCE)
    assert rc == -3                      # B2J_E_UNSUPPORTED (the gate: "missing huffman table")
