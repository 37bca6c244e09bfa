"""Synthetic baseline-JPEG inputs for the parity tests and the bench (SURVEY.md 8d recipe).

Pixels: per channel 128 + sum of 6 amp*sin(fx*x+p1)*cos(fy*y+p2) + N(0,6) noise, seeded with
numpy.random.RandomState(seed). Encoded with Pillow (libjpeg-turbo): baseline, standard or
optimised Huffman tables, chosen subsampling, optional restart interval in MCUs.
"""
import io

import numpy as np
from PIL import Image

SUBSAMPLING = {"444": 0, "422": 1, "420": 2}


def synth_pixels(width, height, seed):
    rng = np.random.RandomState(seed)
    x = np.arange(width, dtype=np.float32)
    y = np.arange(height, dtype=np.float32)
    img = np.empty((height, width, 3), dtype=np.float32)
    for c in range(3):
        acc = np.full((height, width), 128.0, dtype=np.float32)
        for _ in range(6):
            fx, fy = rng.uniform(0.002, 0.08, size=2)
            p1, p2 = rng.uniform(0, 2 * np.pi, size=2)
            amp = rng.uniform(10, 40)
            # amp*sin(fx*x+p1)*cos(fy*y+p2) over the grid == outer product of two 1-D waves
            acc += np.outer(np.cos(fy * y + p2).astype(np.float32), (amp * np.sin(fx * x + p1)).astype(np.float32))
        acc += rng.normal(0, 6, size=(height, width)).astype(np.float32)
        img[:, :, c] = acc
    return np.clip(img, 0, 255).astype(np.uint8)


def encode_jpeg(pixels, quality=90, subsampling="420", restart_mcus=0, optimize=False):
    """pixels: uint8 [H,W,3] RGB. restart_mcus: DRI in MCUs (0 = no restart markers)."""
    im = Image.fromarray(pixels, "RGB")
    buf = io.BytesIO()
    kw = dict(format="JPEG", quality=quality, subsampling=SUBSAMPLING[subsampling], optimize=optimize)
    if restart_mcus:
        kw["restart_marker_blocks"] = int(restart_mcus)
    im.save(buf, **kw)
    return buf.getvalue()


def synth_jpeg(width, height, seed, quality=90, subsampling="420", restart_mcus=0, optimize=False):
    return encode_jpeg(synth_pixels(width, height, seed), quality, subsampling, restart_mcus, optimize)


# BASELINE.json configs (index -> parameters); config 0 is the reference's own fixture.
CONFIGS = {
    1: dict(width=1920, height=1080, quality=90, subsampling="420", restart_mcus=16, count=256),
    2: dict(width=3840, height=2160, quality=95, subsampling="444", restart_mcus=0, count=64),
    3: dict(width=7680, height=4320, quality=85, subsampling="422", restart_mcus=0, count=1),
    4: dict(width=500, height=375, quality=75, subsampling="420", restart_mcus=0, count=8192),
}


def config_jpeg(cfg, index, seed0=1234):
    c = CONFIGS[cfg]
    return synth_jpeg(c["width"], c["height"], seed0 + index, c["quality"], c["subsampling"], c["restart_mcus"])


def _config_job(args):
    cfg, index, seed0 = args
    return config_jpeg(cfg, index, seed0)


def config_batch(cfg, count, first=0, seed0=1234, workers=None):
    """`count` distinct JPEGs of BASELINE config `cfg`, deterministic per index. Generated on a
    thread pool (numpy and Pillow's encoder release the GIL), so it is safe to call from a process
    that already holds a CUDA context -- no fork."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    workers = workers or min(os.cpu_count() or 1, 32)
    jobs = [(cfg, first + i, seed0) for i in range(count)]
    if workers <= 1 or count < 4:
        return [_config_job(j) for j in jobs]
    with ThreadPoolExecutor(workers) as pool:
        return list(pool.map(_config_job, jobs))
