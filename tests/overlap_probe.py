"""Probe (not a pytest file): two batches in flight on two streams vs. one after the other.
usage: python tests/overlap_probe.py [n_images] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import ocljpegdecoder_b200 as b2j
import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
files = synth.config_batch(1, n)
dec = b2j.Decoder(0)
A = dec.batch(files)
B = dec.batch(files)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
p1, p2 = s1.cuda_stream, s2.cuda_stream
A.upload(p1)
B.upload(p2)
torch.cuda.synchronize()


def run(pairs, sa, sb):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(s1)
    s2.wait_event(e0)
    for _ in range(pairs):
        A.decode(sa)
        B.decode(sb)
    done2 = torch.cuda.Event()
    done2.record(s2)
    s1.wait_event(done2)
    e1.record(s1)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * pairs)


for _ in range(2):
    run(3, p1, p1)
print("serial     (both batches on one stream): %.4f ms per step" % run(steps, p1, p1))
for _ in range(2):
    run(3, p1, p2)
print("overlapped (one stream per batch)      : %.4f ms per step" % run(steps, p1, p2))
assert not A.status().any() and not B.status().any()
