"""TEST INFRASTRUCTURE: SHA-256 digests of the reference's coefficient tap and pixels for images of a BASELINE config,
computed on a pool of host processes (spawn: safe next to a live CUDA context). Each worker regenerates its image from
(config, index), so nothing large is pickled. The unmodified reference (oracle/_ref) is used where it decodes the file;
where it aborts (its lost-RSTn defect) or was not built, the oracle port in non-strict mode."""
import hashlib
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

_S = {}


def _job(arg):
    import synth
    from oracle import Oracle, Reference, reference_available
    cfg, index = arg
    if "orc" not in _S:
        _S["orc"] = Oracle()
        _S["ref"] = Reference() if reference_available() else None
    data = synth.config_jpeg(cfg, index)
    kind, coef, bgra = None, None, None
    if _S["ref"] is not None:
        ok, _, coef, bgra, _ = _S["ref"].decode(data, skip_gate=(cfg == 3), want_pixels=True)
        kind = "reference" if ok else None
    if kind is None:
        _S["orc"].set_strict(False)
        rc, _, coef, bgra = _S["orc"].decode(data)
        _S["orc"].set_strict(True)
        if rc != 0:
            return ("failed", "", "")
        kind = "oracle"
    return (kind, hashlib.sha256(coef.tobytes()).hexdigest(), hashlib.sha256(bgra.tobytes()).hexdigest())


def config_digests(cfg, indices, workers=None):
    """[(kind, coef sha256, pixel sha256)] for the images `indices` of BASELINE config `cfg`."""
    workers = workers or min(os.cpu_count() or 1, 32)
    with mp.get_context("spawn").Pool(workers) as pool:
        return pool.map(_job, [(cfg, int(i)) for i in indices], chunksize=max(1, len(indices) // (workers * 4)))
