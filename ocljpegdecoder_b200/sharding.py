"""How a batch of independent images is split over ranks (one process per GPU).

Images are independent units: rank r of W decodes a contiguous slice of the image
list; no data-path collective exists. Only scalar timings/counters are reduced across ranks."""


def shard_range(n_items, rank, world):
    """Contiguous, balanced slice [lo, hi) of range(n_items) for `rank` (sizes differ by at most 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_bytes(sizes, world):
    """Greedy contiguous split of items with byte sizes `sizes` into `world` slices of similar total
    bytes (compressed size is the work proxy for the entropy decoder). Returns a list of (lo, hi)."""
    total = sum(sizes)
    out, lo, acc = [], 0, 0
    for r in range(world):
        target = total * (r + 1) / world
        hi = lo
        while hi < len(sizes) and (acc + sizes[hi] <= target or hi == lo) and len(sizes) - hi > world - r - 1:
            acc += sizes[hi]
            hi += 1
        if r == world - 1:
            hi = len(sizes)
        out.append((lo, hi))
        lo = hi
    return out
