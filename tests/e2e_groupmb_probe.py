"""Probe (not a pytest file): b2j_decode_host_ex wall time for several byte bounds of a pipeline group.
usage: python tests/e2e_groupmb_probe.py config n_images"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ocljpegdecoder_b200 as b2j, synth
cfg = int(sys.argv[1]); n = int(sys.argv[2])
files = synth.config_batch(cfg, n)
c = synth.CONFIGS[cfg]
npix = c["width"] * c["height"]
dec = b2j.Decoder(0)
pinned = b2j.PinnedBuffer(n * npix * 4)
outs = [pinned.address + i * npix * 4 for i in range(n)]
for mb in (6, 12, 24, 48, 96, 1024):
    args = b2j.HostArgs(files, outs=outs, group_mb=mb)
    ts = []
    for rep in range(5):
        t0 = time.perf_counter(); _, st = dec.decode_host_args(args); ts.append(1e3 * (time.perf_counter() - t0))
    assert not st.any()
    print("config %d group_mb %4d: min %.2f ms  median %.2f ms" % (cfg, mb, min(ts[1:]), sorted(ts[1:])[2]), flush=True)
dec.close(); pinned.close()
