"""Probe (not a pytest file): b2j_decode_host wall time for several group sizes, with and without the ramp."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import ocljpegdecoder_b200 as b2j, synth
files = synth.config_batch(1, 256)
dec = b2j.Decoder(0)
outs_t = [torch.empty((1080, 1920, 4), dtype=torch.uint8, pin_memory=True) for _ in range(256)]
outs = [o.numpy() for o in outs_t]
for ramp in ("0", "1"):
    for group in ("8", "16", "32", "64"):
        os.environ["B2J_HOST_RAMP"] = ramp
        os.environ["B2J_HOST_GROUP"] = group
        ts = []
        for rep in range(6):
            t0 = time.perf_counter(); dec.decode_host(files, outs); ts.append(1e3 * (time.perf_counter() - t0))
        print("ramp %s group %2s: min %.2f ms  median %.2f ms" % (ramp, group, min(ts[1:]), sorted(ts[1:])[2]), flush=True)
