#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native baseline-JPEG decode path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, host cores

Metric (BASELINE.json): MPix/s decoded on the 1080p 4:2:0 q90 batch (configs[1]: 256 synthetic
1920x1080 baseline JPEGs, restart interval 16 MCUs) -- per GPU; with N GPUs every rank decodes
its own 256 distinct images (sharded by image, no collective on the data path => weak scaling).
A "step" is one pass of the hot path (pre-pass -> Huffman -> IDCT/colour) over the whole batch.

  value      whole-job MPix/s with the compressed batch already resident in HBM; K steps timed with
             CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
  e2e        same metric through b2j_decode_host(): host JPEG bytes in, host (pinned) BGRA out; header
             parsing, staging, H2D, decode and D2H are all inside the timed region.
  roofline   dominant kernel: algorithmic bytes per launch / its mean launch time (CUDA events inside
             the timed region) against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference CPU path (oracle/_ref when present, else the oracle port) on the host
             cores of this box, bounded sample. The only place oracle/ is executed by this file.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "MPix/s decoded (1080p 4:2:0 q90 batch)"
WORKLOAD = "256 synthetic 1920x1080 4:2:0 baseline q90 JPEGs, restart interval 16 MCUs (BASELINE configs[1])"
CFG = 1            # index into tests/synth.py CONFIGS
BATCH = 256
FALLBACK_HBM_GBS = 6650.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
# CPU side (reference arm / cpu_baseline): runs before any CUDA context exists (fork-safe).
_W = {}


def _cpu_worker_init():
    from oracle import Oracle, Reference, reference_available
    _W["ref"] = Reference() if reference_available() else None
    _W["orc"] = Oracle()


def _cpu_decode_one(data):
    """Returns (pixels, huffman seconds, idct+colour seconds); pixels == 0 when the reference fails."""
    if _W["ref"] is not None:
        ok, info, _, _, ts = _W["ref"].decode(data, skip_gate=False, want_pixels=True)
        return (info.width * info.height if ok else 0), ts[0], ts[1]
    t0 = time.perf_counter()
    rc, img, _, _ = _W["orc"].decode(data)   # strict mode: fails where the reference would
    return (img.width * img.height if rc == 0 else 0), time.perf_counter() - t0, 0.0


def cpu_pool(cores):
    import multiprocessing as mp
    return mp.get_context("fork").Pool(cores, initializer=_cpu_worker_init)


def cpu_kind():
    from oracle import reference_available
    return "reference" if reference_available() else "port"


def run_cpu_sample(pool, files):
    """Decodes `files` once over all workers; returns (MPix/s wall, wall s, sum huffman s, sum idct+colour s)."""
    t0 = time.perf_counter()
    res = pool.map(_cpu_decode_one, files, chunksize=1)
    dt = time.perf_counter() - t0
    pix = sum(r[0] for r in res)
    return pix / dt / 1e6, dt, sum(r[1] for r in res), sum(r[2] for r in res)


def screen_cpu_files(pool, files, want):
    """The reference loses an RSTn whose FF is the last byte of one of its 2 KiB reads
    (decoder.cpp:118-131) and then aborts the image: about 1 in 5 of the 1080p RI=16
    images. Such images are left out of the CPU timing (an aborted decode would flatter it).
    Returns (usable files, number of candidates screened, number the reference failed on)."""
    good, seen, bad = [], 0, 0
    for k in range(0, len(files), 32):
        part = files[k:k + 32]
        res = pool.map(_cpu_decode_one, part, chunksize=1)
        for f, r in zip(part, res):
            seen += 1
            if r[0]:
                good.append(f)
            else:
                bad += 1
        if len(good) >= want:
            break
    return good[:want], seen, bad


def reference_arm(args, rank, world):
    if rank != 0:
        return 0
    import synth
    import __graft_entry__ as ge
    ge.build()
    cores = os.cpu_count() or 1
    per_step = max(cores * 2, 16)                       # bounded sample of the 256-image workload
    pool = cpu_pool(cores)
    files, seen, bad = screen_cpu_files(pool, synth.config_batch(CFG, min(BATCH, per_step * 2)), per_step)
    per_step = len(files)
    for _ in range(args.warmup):
        run_cpu_sample(pool, files[:cores])
    t_total, pix_total = 0.0, 0
    for _ in range(args.steps):
        mp, dt, _, _ = run_cpu_sample(pool, files)
        t_total += dt
        pix_total += mp * dt * 1e6
    pool.close()
    value = pix_total / t_total / 1e6
    sample = ("%d of the 256 images per step, one process per host core (the reference is single-threaded); "
              "%d of %d candidates dropped because the reference aborts on them" % (per_step, bad, seen))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "MPix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t_total / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": round(value, 3), "unit": "MPix/s", "cores": cores, "kind": cpu_kind(), "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:   # NVML missing: report nothing rather than guess
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag.is_set():
            try:
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        busy = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        return {"sm_mhz": int(statistics.median(busy)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r01_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            return int(json.load(f)[kernel])
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b2j", choices=["b2j", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the cpu_baseline sample (0 = 8 per core)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b2j" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import numpy as np
    import synth
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()

    # ---- inputs (distinct images per rank) and the CPU baseline: both before CUDA is touched
    t0 = time.time()
    files = synth.config_batch(CFG, BATCH, first=rank * BATCH)
    log("[rank %d] generated %d JPEGs, %.1f MB, %.1f s" % (rank, len(files), sum(map(len, files)) / 1e6, time.time() - t0))
    cpu = None
    if rank == 0 and world == 1:
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_sample or min(BATCH, cores * 8)
        pool = cpu_pool(cores)
        cpu_files, seen, bad = screen_cpu_files(pool, files, n_cpu)   # also warms the workers
        n_cpu = len(cpu_files)
        mp_all, dt_all, th, tm = run_cpu_sample(pool, cpu_files)
        pool.close()
        one_core = n_cpu * 1920 * 1080 / (th + tm) / 1e6 if (th + tm) > 0 else None
        cpu = {"value": round(mp_all, 3), "unit": "MPix/s", "cores": cores, "kind": cpu_kind(),
               "sample": "%d of the %d images, one process per host core, %.1f s wall; the reference aborts on %d of %d "
                         "screened images (lost RSTn at a 2 KiB read boundary), those are excluded" % (n_cpu, BATCH, dt_all, bad, seen),
               "one_core_mpix_s": round(one_core, 3) if one_core else None,
               "huffman_share": round(th / (th + tm), 3) if (th + tm) > 0 else None}
        log("[cpu] %s" % cpu)

    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    if rank != 0:
        ge.build()
    import ocljpegdecoder_b200 as b2j

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dec = b2j.Decoder(local_rank)
    batch = dec.batch(files)
    info = batch.info()
    batch.upload()
    batch.sync()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-resident timing: W warm-up steps, then exactly K steps
    batch.decode_steps(args.warmup)
    barrier()
    per, total_ms = batch.decode_steps(args.steps)
    barrier()
    st = batch.status()
    assert not st.any(), "decode status not clean: %s" % st[st != 0][:8]
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    pix_per_step = info.total_pixels * world
    value = pix_per_step / (ms_per_step * 1e-3) / 1e6

    stage = {k: statistics.mean(getattr(p, k) for p in per) for k in ("prepass_ms", "huffman_ms", "idct_ms", "total_ms")}

    # ---- end to end through the public host API: JPEG bytes in host memory -> BGRA in pinned host memory
    outs_t = [torch.empty((1080, 1920, 4), dtype=torch.uint8, pin_memory=True) for _ in range(BATCH)]
    outs = [o.numpy() for o in outs_t]
    dec.decode_host(files, outs)                     # warm-up (allocations, page faults)
    dec.decode_host(files, outs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        _, st2 = dec.decode_host(files, outs)
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    assert not st2.any()
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = pix_per_step / float(t.item()) / 1e6

    sampler.stop_flag.set()
    sampler.join(timeout=2)

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel and its algorithmic bytes per launch (DESIGN.md "Roofline arithmetic"):
        #   Huffman kernel : scan bytes read + int16 coefficient plane written  = C + 128*B
        #   IDCT/colour    : coefficient plane read + BGRA written              = 128*B + 4*W*H
        kb = {"huffman_ms": ("k_huff_decode", info.scan_bytes + info.coef_plane_bytes),
              "idct_ms": ("k_idct_csc", info.coef_plane_bytes + info.pixel_bytes)}
        dom = max(kb, key=lambda k: stage[k])
        achieved = kb[dom][1] / (stage[dom] * 1e-3) / 1e9
        pipe = info.algorithmic_bytes / (stage["total_ms"] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "MPix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_gpu": BATCH, "scan_bytes_per_gpu": info.scan_bytes,
                       "blocks_per_gpu": info.total_blocks, "algorithmic_bytes_per_step_per_gpu": info.algorithmic_bytes,
                       "l2": "no explicit flush: every step streams %.2f GB per GPU (>> 126 MB L2)" % (info.algorithmic_bytes / 1e9),
                       "parallelism": "images sharded by rank, no collective"},
            "roofline": {"bound": "hbm", "kernel": kb[dom][0], "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": measured_traffic(kb[dom][0]), "peak_source": peak_src,
                         "bytes_per_launch": kb[dom][1], "launch_ms": round(stage[dom], 4),
                         "pipeline": {"achieved": round(pipe, 1), "frac": round(pipe / peak, 4),
                                      "note": "all kernels of a step (pre-pass, Huffman, IDCT+colour): (C + 2*128*B + 4*W*H) / step time, SURVEY.md 8(d)"},
                         "stage_ms": {k: round(v, 4) for k, v in stage.items()},
                         "huffman_gbit_s": round(info.scan_bytes * 8 / (stage["huffman_ms"] * 1e-3) / 1e9, 1)},
            "cpu_baseline": cpu,
            "e2e": {"value": round(e2e_value, 1), "unit": "MPix/s", "h2d_bytes_per_step": info.h2d_bytes * world,
                    "d2h_bytes_per_step": info.pixel_bytes * world, "ms_per_step": round(float(t.item()) * 1e3, 3),
                    "api": "b2j_decode_host: parse + stage + H2D + decode + D2H into pinned host buffers"},
            "gpu_launches": info.kernel_launches * args.steps,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    batch.close()
    dec.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
