"""Short device-resident run for ncu (not a pytest file): config-2 batch, a few decode steps.
usage: python tests/prof_run.py [n_images] [steps] [config]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ocljpegdecoder_b200 as b2j
import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 1
files = synth.config_batch(cfg, n)
dec = b2j.Decoder(0)
batch = dec.batch(files)
batch.upload()
per, total = batch.decode_steps(steps)
st = batch.status()
assert not st.any(), st
print("sync stats", batch.sync_stats().tolist(), "blocks", batch.info().total_blocks)
print("steps %d total %.3f ms; last step: prepass %.3f huffman %.3f idct %.3f" % (
    steps, total, per[steps - 1].prepass_ms, per[steps - 1].huffman_ms, per[steps - 1].idct_ms))
