"""Multi-GPU plumbing of the path for the process-per-GPU launch (torchrun): the image tile space shards trivially.

The partition itself lives in the library: `pbrt_gpu_render_tiles_device(scene, params, rank, world, ...)` renders part `rank` of the
frame's 16x16 tiles, dealt round robin in the Morton order the reference's BlockQueue hands tiles to its threads in
(src/blockqueue/mod.rs:33-36) -- and `pbrt_gpu_render_multi` does the whole frame, reduce included, from one process over N devices.
What is left for the per-process launch is the one collective: every rank renders into a full-size film that is zero elsewhere and
ONE reduce(sum) merges the films on rank 0 -- a sum, not a gather, so filter footprints that cross a tile border merge exactly like
Film::merge_film_tile's `+=` (film.rs:362-367).  `band` is the contiguous-rows partition round 1 used (kept for the CPU tests and as the
unbalanced baseline the tile interleave is measured against).
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
