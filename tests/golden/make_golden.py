"""Generates the golden vectors under tests/golden/ (run in the build container, where the
unmodified reference can be compiled: oracle/build_ref.sh).

For each small synthetic JPEG (kept as a file, so a different libjpeg build cannot change it)
the UNMODIFIED reference CPU path produces the coefficient tap and the pixel tap; their SHA-256
and a few spot values go to golden.json. The fixture of the reference repository itself
(test/JPEG_example_JPG_RIP_050.jpg) is included with the hashes SURVEY.md 8(c) records.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

import synth
from oracle import Reference

CASES = [
    # name, width, height, subsampling, quality, restart interval (MCUs), optimised tables, seed
    ("g444_q75", 64, 48, "444", 75, 0, False, 11),
    ("g420_odd_q90", 67, 45, "420", 90, 0, False, 12),
    ("g420_ri1_q90", 67, 45, "420", 90, 1, False, 13),
    ("g420_ri3_opt_q50", 200, 120, "420", 50, 3, True, 14),
    ("g422_q85", 130, 70, "422", 85, 0, False, 15),
    ("g422_ri2_opt_q85", 130, 70, "422", 85, 2, True, 16),
    ("g444_ri5_opt_q95", 133, 81, "444", 95, 5, True, 17),
    ("g420_ri11_q100", 160, 96, "420", 100, 11, False, 18),
    ("g444_q10_tiny", 8, 8, "444", 10, 0, False, 19),
    ("g420_ri9_many", 320, 64, "420", 80, 2, False, 20),   # 80 intervals: RSTn wraps mod 8 ten times
]


def main():
    ref = Reference()
    out = {}
    with open(os.path.join(HERE, "JPEG_example_JPG_RIP_050.jpg"), "rb") as f:
        fixture = f.read()
    entries = [("JPEG_example_JPG_RIP_050", fixture, False)]
    for name, w, h, ss, q, ri, opt, seed in CASES:
        data = synth.synth_jpeg(w, h, seed, q, ss, ri, opt)
        with open(os.path.join(HERE, name + ".jpg"), "wb") as f:
            f.write(data)
        entries.append((name, data, ss == "422"))
    for name, data, skip_gate in entries:
        ok, info, coef, bgra, _ = ref.decode(data, skip_gate=skip_gate)
        assert ok, name
        out[name] = {
            "file_sha256": hashlib.sha256(data).hexdigest(),
            "width": info.width, "height": info.height, "blk_count": info.blk_count,
            "sampling": list(info.sampling), "restart_interval": info.restart_interval,
            "needs_extended_gate": bool(skip_gate),
            "coef_sha256": hashlib.sha256(coef.tobytes()).hexdigest(),
            "pixel_sha256": hashlib.sha256(bgra.tobytes()).hexdigest(),
            "coef_sum_abs": int(np.abs(coef.astype(np.int64)).sum()),
            "pixel_sum": int(bgra.astype(np.int64).sum()),
        }
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote %d golden entries" % len(out))


if __name__ == "__main__":
    main()
