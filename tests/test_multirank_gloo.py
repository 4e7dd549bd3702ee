"""World-size-2 test of the N > 1 host logic on CPU (gloo): band partition + one reduce(sum) of the film
reproduces the single-rank film.  The renderer stand-in here is the oracle (test infrastructure); on GPUs
bench.py uses pbrt_gpu_render_device + NCCL with the same two helpers."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rs_pbrt_b200.multigpu import band, reduce_film

ROOT = Path(__file__).resolve().parent.parent


def test_band_partition_covers_rect_exactly():
    for rect in ([0, 0, 400, 400], [3, 5, 100, 77], [0, 0, 1920, 1080], [0, 0, 8, 8]):
        for world in (1, 2, 3, 4, 8):
            rows = []
            for r in range(world):
                b = band(rect, r, world)
                assert b[0] == rect[0] and b[2] == rect[2] and rect[1] <= b[1] <= b[3] <= rect[3]
                rows += list(range(b[1], b[3]))
            assert rows == list(range(rect[1], rect[3]))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib
    from rs_pbrt_b200 import scenes

    h = scenes.cornell_box(xres=40, yres=40, spp=4, filter="gaussian", xwidth=1.5, ywidth=1.5)
    rp = h.params.contents
    osc = oracle_lib.OracleScene(h.desc)
    film, _, _ = osc.render(h.params, rect=band(list(rp.sample_bounds), rank, world), n_threads=1)
    t = torch.from_numpy(film)
    reduce_film(t, dist)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _tile_worker(rank, world, port, out_path):
    """The N > 1 path as bench.py runs it, on CPU: every rank renders ITS TILE SHARE with the product's own
    pbrt_gpu_render_tiles_device -- through the kernel-emulation build of the library (tests/emu) -- and gloo sums the films."""
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C

    from rs_pbrt_b200 import GpuScene, _abi, scenes

    E = _abi.bind(C.CDLL(str(ROOT / "tests" / "emu" / "_build" / "librs_pbrt_b200_emu.so")))
    h = scenes.cornell_box(xres=40, yres=24, spp=2, filter="gaussian", xwidth=1.5, ywidth=1.5)
    g = GpuScene(h.desc, 0, lib=E)
    film = np.zeros((24, 40, 4), np.float32)
    st = g.render_tiles_device(h.params, film.ctypes.data, rank, world)
    g.close()
    t = torch.from_numpy(film)
    rays = torch.tensor([float(st["rays"])], dtype=torch.float64)
    reduce_film(t, dist)
    dist.reduce(rays, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, np.concatenate([t.numpy().reshape(-1), rays.numpy().astype(np.float32)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_tile_shares_reproduce_the_frame(tmp_path, oracle):
    import shutil

    import pytest
    if shutil.which("g++") is None:
        pytest.skip("no g++ for the kernel-emulation build")
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu

    build_emu.build()
    from rs_pbrt_b200 import scenes

    out = tmp_path / "film_tiles.npy"
    port = 31500 + os.getpid() % 2000
    mp.spawn(_tile_worker, args=(2, port, str(out)), nprocs=2, join=True)
    got = np.load(out)
    h = scenes.cornell_box(xres=40, yres=24, spp=2, filter="gaussian", xwidth=1.5, ywidth=1.5)
    ref, _, so = oracle.OracleScene(h.desc).render(h.params, n_threads=2)
    assert int(got[-1]) == so["rays"]
    assert np.allclose(got[:-1].reshape(24, 40, 4), ref, rtol=1e-5, atol=1e-6)


def test_two_ranks_reproduce_the_single_rank_film(tmp_path, oracle):
    from rs_pbrt_b200 import scenes

    out = tmp_path / "film.npy"
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    got = np.load(out)
    h = scenes.cornell_box(xres=40, yres=40, spp=4, filter="gaussian", xwidth=1.5, ywidth=1.5)
    ref, _, _ = oracle.OracleScene(h.desc).render(h.params, n_threads=1)
    # the wide filter makes the two bands overlap by several rows: only a SUM merges them correctly
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)
    assert np.count_nonzero(ref[..., 3]) == 40 * 40
