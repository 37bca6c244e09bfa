"""Probe (not a pytest file): stage times of config 2 for every output format."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ocljpegdecoder_b200 as b2j, synth
files = synth.config_batch(1, 256)
dec = b2j.Decoder(0)
batch = dec.batch(files)
batch.upload()
for name, fmt in (("BGRA", b2j.OUT_BGRA), ("RGB24", b2j.OUT_RGB24), ("planar RGB", b2j.OUT_RGB_PLANAR), ("BGRA", b2j.OUT_BGRA)):
    batch.set_output_format(fmt)
    per, total = batch.decode_steps(20)
    assert not batch.status().any()
    print("%-10s steps 20 total %.3f ms; last step: prepass %.3f huffman %.3f idct %.3f" % (name, total, per[19].prepass_ms, per[19].huffman_ms, per[19].idct_ms))
