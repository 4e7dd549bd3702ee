"""Multi-GPU plumbing of the path: the image tile space shards trivially.

The scene is replicated; rank r renders a contiguous band of 16-pixel tile rows into a full-size film
that is zero elsewhere, and ONE reduce(sum) merges the films on rank 0 -- a sum, not a gather, so filter
footprints that cross a band border merge exactly like Film::merge_film_tile's `+=` (film.rs:362-367).
"""


def band(rect, rank, world, tile=16):
    """Pixel rectangle {x0,y0,x1,y1} of `rank`'s band of tile rows (tile = 16 as integrator.rs:75)."""
    x0, y0, x1, y1 = (int(v) for v in rect)
    rows = max(0, y1 - y0)
    tiles = (rows + tile - 1) // tile
    t0 = tiles * rank // world
    t1 = tiles * (rank + 1) // world
    return [x0, min(y0 + tile * t0, y1), x1, min(y0 + tile * t1, y1)]


def reduce_film(film, dist, dst=0):
    """Sum the per-rank films onto `dst` (NCCL on GPUs, gloo in the CPU tests)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film
