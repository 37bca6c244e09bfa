"""Where the end-to-end time of b2j_decode_host goes (manual GPU check, not a pytest file)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import ocljpegdecoder_b200 as b2j, synth
files = synth.config_batch(1, 256)
dec = b2j.Decoder(0)
outs_t = [torch.empty((1080, 1920, 4), dtype=torch.uint8, pin_memory=True) for _ in range(256)]
outs = [o.numpy() for o in outs_t]
for rep in range(3):
    t0 = time.perf_counter(); batch = dec.batch(files); t1 = time.perf_counter()
    batch.upload(); batch.sync(); t2 = time.perf_counter()
    batch.decode(); batch.sync(); t3 = time.perf_counter()
    batch.read_all_pixels(outs); t4 = time.perf_counter()
    st = batch.status(); batch.close(); t5 = time.perf_counter()
    print("create %.2f ms  upload %.2f  decode %.2f  d2h %.2f  status+close %.2f  total %.2f" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)))
for rep in range(3):
    t0 = time.perf_counter(); dec.decode_host(files, outs); print("decode_host %.2f ms" % (1e3 * (time.perf_counter() - t0)))
