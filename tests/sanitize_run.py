"""Small mixed batch for compute-sanitizer (manual GPU check, not a pytest file)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ocljpegdecoder_b200 as b2j, synth
files = [open(os.path.join(ROOT, "tests", "golden", "JPEG_example_JPG_RIP_050.jpg"), "rb").read()]
for i, (w, h, ss, q, ri) in enumerate([(67, 45, "420", 90, 1), (200, 120, "444", 50, 3), (130, 70, "422", 85, 0), (320, 240, "420", 90, 16), (333, 211, "444", 95, 0), (31, 257, "420", 100, 0)]):
    files.append(synth.synth_jpeg(w, h, 50 + i, q, ss, ri))
dec = b2j.Decoder(0)
batch = dec.batch(files)
batch.upload(); batch.decode()
print("status", batch.status().tolist(), "sync", batch.sync_stats().tolist())
c = batch.coefs(0); p = batch.pixels(3)
outs, st = dec.decode_host(files)
print("host status", st.tolist(), c.shape, p.shape)
