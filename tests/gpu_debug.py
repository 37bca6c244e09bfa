"""Manual GPU check with verbose mismatch reporting (not a pytest file)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import __graft_entry__ as ge

ge.build()
import ocljpegdecoder_b200 as b2j
from oracle import Oracle
import synth


def report(name, got, want):
    if got.shape != want.shape:
        print("  %s: SHAPE %s vs %s" % (name, got.shape, want.shape))
        return False
    bad = np.argwhere(got != want)
    if len(bad) == 0:
        print("  %s: exact (%d values)" % (name, got.size))
        return True
    print("  %s: %d / %d differ; first at %s got %s want %s; max|d| %d" % (
        name, len(bad), got.size, bad[0], got[tuple(bad[0])], want[tuple(bad[0])],
        np.abs(got.astype(np.int64) - want.astype(np.int64)).max()))
    if name == "coef":
        blks = np.unique(bad[:, 0])
        print("    bad blocks: %d, first %s last %s" % (len(blks), blks[:8], blks[-3:]))
    return False


def main():
    orc = Oracle()
    with open(os.path.join(ROOT, "tests", "golden", "JPEG_example_JPG_RIP_050.jpg"), "rb") as f:
        fixture = f.read()
    cases = [("fixture", fixture)]
    specs = [(64, 48, "444", 75, 0), (67, 45, "420", 90, 0), (67, 45, "420", 90, 1), (200, 120, "420", 50, 3),
             (130, 70, "422", 85, 0), (130, 70, "422", 85, 2), (640, 480, "420", 90, 16), (640, 480, "444", 95, 7),
             (333, 211, "444", 95, 5), (1920, 1080, "420", 90, 16), (16, 16, "420", 100, 0), (8, 8, "444", 10, 0)]
    for i, (w, h, ss, q, ri) in enumerate(specs):
        cases.append(("%dx%d_%s_q%d_ri%d" % (w, h, ss, q, ri), synth.synth_jpeg(w, h, 100 + i, q, ss, ri)))
    dec = b2j.Decoder(0)
    allok = True
    # one batch with everything (mixed layouts in one launch), then each alone
    files = [c[1] for c in cases]
    t0 = time.time()
    batch = dec.batch(files)
    batch.upload()
    times = batch.decode_timed()
    print("batch of %d: prepass %.3f ms, huffman %.3f ms, idct %.3f ms (create+upload+decode wall %.3f s)" % (
        len(files), times.prepass_ms, times.huffman_ms, times.idct_ms, time.time() - t0))
    st = batch.status()
    print("status:", st)
    for i, (name, f) in enumerate(cases):
        rc, img, coef, bgra = orc.decode(f)
        print(name, "rc", rc, "blocks", img.blk_count)
        allok &= report("coef", batch.coefs(i), coef)
        allok &= report("pix", batch.pixels(i), bgra)
    batch.close()
    print("ALL OK" if allok else "MISMATCHES")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
