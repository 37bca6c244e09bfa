"""Probe (not a pytest file): b2j_decode_host_ex wall time for several group sizes and host-thread counts.
usage: python tests/e2e_probe.py config n_images"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ocljpegdecoder_b200 as b2j, synth
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
files = synth.config_batch(cfg, n)
c = synth.CONFIGS[cfg]
npix = c["width"] * c["height"]
dec = b2j.Decoder(0)
pinned = b2j.PinnedBuffer(n * npix * 4)
outs = [pinned.address + i * npix * 4 for i in range(n)]
print("config %d, %d images, %d host cores; PCIe floor of the BGRA download at 56 GB/s: %.1f ms" % (cfg, n, os.cpu_count(), n * npix * 4 / 56e9 * 1e3))
for threads in (1, 2, 4, 8, 16):
    for group in (16, 32, 64, 128, 256):
        ts = []
        for rep in range(4):
            t0 = time.perf_counter()
            _, st = dec.decode_host_ex(files, outs=outs, n_threads=threads, group=group)
            ts.append(1e3 * (time.perf_counter() - t0))
        assert not st.any()
        print("threads %2d group %3d: min %.2f ms  median %.2f ms" % (threads, group, min(ts[1:]), sorted(ts[1:])[1]), flush=True)
dec.close()
pinned.close()
