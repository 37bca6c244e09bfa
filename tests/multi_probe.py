"""Probe (not a pytest file): b2j_decode_host_multi -- ONE process, one context and one host thread per GPU (C ABI).
usage: python tests/multi_probe.py config n_images [steps]      (all GPUs of the box; also run with CUDA_VISIBLE_DEVICES)
Prints e2e MPix/s (host JPEG bytes -> pinned host pixels) for BGRA and RGB24 output with 1, 2, 4, ... GPUs."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ocljpegdecoder_b200 as b2j
import synth

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
files = synth.config_batch(cfg, n)
c = synth.CONFIGS[cfg]
npix = c["width"] * c["height"]
ngpu = b2j.load_library().b2j_device_count()
decs = [b2j.Decoder(k) for k in range(ngpu)]
pinned = b2j.PinnedBuffer(n * npix * 4)
print("config %d: %d images, %.1f MB compressed, %d GPUs, %d host cores" % (cfg, n, sum(map(len, files)) / 1e6, ngpu, os.cpu_count()))
for fmt, name, bpp in ((b2j.OUT_BGRA, "BGRA", 4), (b2j.OUT_RGB24, "RGB24", 3)):
    outs = [pinned.address + i * npix * bpp for i in range(n)]
    g = 1
    while g <= ngpu:
        for _ in range(2):
            b2j.decode_host_multi(decs[:g], files, outs=outs, out_format=fmt)      # warm-up: pools grow to what the pipeline holds in flight
        t0 = time.perf_counter()
        for _ in range(steps):
            _, st = b2j.decode_host_multi(decs[:g], files, outs=outs, out_format=fmt)
        dt = (time.perf_counter() - t0) / steps
        assert not st.any()
        print("%-5s %d GPU(s): %8.2f ms per batch, %9.1f MPix/s, %.1f GB/s of pixels to the host" % (name, g, dt * 1e3, n * npix / dt / 1e6, n * npix * bpp / dt / 1e9), flush=True)
        g *= 2
first = np.ctypeslib.as_array  # keep numpy imported for the pinned view
for d in decs:
    d.close()
pinned.close()
