"""CPU: host emulation of the self-synchronising entropy path (tests/native/synccheck.cpp).

The walk (walk_stream with the multi-symbol walk tables) and the chunk-wise synchronisation rounds are the same
inline code the kernels compile (ocljpegdecoder_b200/csrc/b2j_sync.h). Here they run lane after lane on the CPU and
every sub-sequence record -- exit state, blocks started, DC sums, first block start -- is compared with a sequential,
table-free, one-symbol-at-a-time walk of the same stream (the reference's scan order, decoder.cpp:221-346)."""
import ctypes
import os

import numpy as np
import pytest

import synth
from conftest import ROOT

GATE_EXTENDED, GATE_GRAY = 1, 4


@pytest.fixture(scope="module")
def synccheck(built):
    L = ctypes.CDLL(os.path.join(ROOT, "tests", "native", "libb2jsync.so"))
    L.b2j_synccheck.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]

    def run(data, force_sweep=False, gate=GATE_EXTENDED):
        stats = np.zeros(10, np.uint64)
        rc = L.b2j_synccheck(data, len(data), gate, 1 if force_sweep else 0, stats.ctypes.data)
        return rc, stats
    return run


CASES = [(500, 375, "420", 75, False), (640, 480, "444", 95, False), (640, 360, "422", 5, False), (257, 129, "420", 100, True),
         (257, 129, "444", 100, False), (1024, 768, "444", 92, True), (333, 777, "420", 60, False), (1280, 720, "422", 85, False),
         (64, 48, "444", 75, False), (8, 8, "444", 10, False), (1, 1, "420", 90, False)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_%s_q%d%s" % (c[0], c[1], c[2], c[3], "_opt" if c[4] else ""))
def test_walk_and_chunk_rounds_agree_with_sequential_walk(synccheck, case):
    w, h, ss, q, opt = case
    data = synth.synth_jpeg(w, h, 4000 + w + q, q, ss, 0, optimize=opt)
    rc, st = synccheck(data)
    assert rc == 0, (rc, st.tolist())
    assert st[9] == 0
    # the sweep is the safety net: with pre-lanes switched off every chunk border must go through it and still agree
    rc, st2 = synccheck(data, force_sweep=True)
    assert rc == 0, (rc, st2.tolist())
    if st2[1] > 1:
        assert st2[5] > 0


def test_walk_tables_take_several_symbols_per_step(synccheck):
    """The point of the walk tables: at photographic qualities most steps cover more than one symbol."""
    data = synth.synth_jpeg(1920, 1080, 77, 95, "444", 0)
    rc, st = synccheck(data)
    assert rc == 0
    steps = int(st[6] + st[7])
    assert st[7] < 0.25 * steps, st.tolist()          # few steps fall back to the one-symbol path
    print("sub-sequences %d, walk-table steps %d, one-symbol steps %d, round-1 lanes %d (missed the checkpoint: %d), sweep %d"
          % (st[0], st[6], st[7], st[2], st[3], st[5]))


def test_grayscale_stream(synccheck):
    import io
    from PIL import Image
    px = synth.synth_pixels(320, 200, 5)[:, :, 1]
    buf = io.BytesIO()
    Image.fromarray(px, "L").save(buf, format="JPEG", quality=85)
    rc, st = synccheck(buf.getvalue(), gate=GATE_EXTENDED | GATE_GRAY)
    assert rc == 0, (rc, st.tolist())
