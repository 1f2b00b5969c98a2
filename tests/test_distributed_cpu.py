"""N > 1 path on CPU: two gloo processes own bands of tile rows (distributed.band_of_rank), each computes ITS tiles with
the CPU oracle (standing in for the GPU engine), and the disjoint output bands are gathered on rank 0
(distributed.gather_bands).  The stitched rasters must equal the single-process result bit for bit: tiles are
self-sufficient, no arithmetic crosses ranks (mode A)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    import golden_inputs
    import toy_models
    from moonsuperresolution_b200.distributed import band_of_rank, gather_bands
    from moonsuperresolution_b200.planner import Plan
    from oracle import tiling as OT
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    case = golden_inputs.CASES["wobble_1100x1300"]
    dem, img = golden_inputs.make_rasters(case)
    h, w, i, s, b, t, nv = case["H"], case["W"], case["I"], case["S"], case["B"], case["T"], case["NV"]
    plan = Plan(h, w, i, s, t, b)
    geo = OT.Geometry(h, w, i, s, t)
    tiles, r0, r1 = band_of_rank(plan, world, rank)
    dem_c, img_c = OT.pad_inputs(dem, img, geo, nv)
    bands = [np.zeros((r1 - r0, w), np.float32), np.zeros((r1 - r0, w), np.float32), np.zeros((r1 - r0, w), np.uint8)]
    for (px, py) in tiles:
        m, sd, g = OT.process_tile(dem_c, img_c, geo, px, py, b, nv, toy_models.wobble)
        rows, cols = plan.tile_window(px, py)
        for k, a in enumerate((m, sd, g)):
            bands[k][py - r0:py - r0 + rows, px:px + cols] = a[:rows, :cols]
    res = gather_bands(bands, r0, h, w, rank, world)
    if rank == 0:
        np.savez(out_path, mean=res[0], std=res[1], good=res[2])
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_band_sharding_gloo(tmp_path, golden, world):
    import hashlib
    import torch.multiprocessing as mp
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    z = np.load(out)
    name = "wobble_1100x1300"
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    # equal to what the reference's own single-process class produced (golden fixture)
    assert sha(z["mean"]) == str(golden[f"{name}/sha_mean"])
    assert sha(z["std"]) == str(golden[f"{name}/sha_std"])
    assert sha(z["good"]) == str(golden[f"{name}/sha_good"])


def _halo_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from moonsuperresolution_b200.distributed import band_of_rank, exchange_halo_rows, rows_read_by_band
    from moonsuperresolution_b200.planner import Plan
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    h, w = 2900, 37
    plan = Plan(h, w, 64, 16, 256, 4)
    full = torch.arange(h * w, dtype=torch.float32).reshape(h, w)
    _, r0, r1 = band_of_rank(plan, world, rank)
    got = exchange_halo_rows(full[r0:r1].clone(), plan, rank, world)
    n0, n1 = rows_read_by_band(plan, world, rank)
    assert n0 <= r0 and n1 >= r1 and (rank == 0 or n0 == r0 - plan.off) and (rank == world - 1 or n1 == r1 + plan.off)
    assert torch.equal(got, full[n0:n1]), f"rank {rank}: halo rows differ"
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_input_halo_exchange_gloo(tmp_path, world):
    """Sharded loading: each rank holds only its own band of raster rows and gets the I - S halo rows from its neighbours
    by point-to-point send / recv (NCCL over NVLink on the GPU box, gloo here)."""
    import torch.multiprocessing as mp
    mp.spawn(_halo_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def _dedup_halo_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from moonsuperresolution_b200.distributed import exchange_halo_rows
    from moonsuperresolution_b200.planner import Plan
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    h, w = 2900, 37
    plan = Plan(h, w, 64, 16, 256, 4)
    full = torch.arange(h * w, dtype=torch.float32).reshape(h, w)
    rows = [plan.dedup_rows(world, r) for r in range(world)]
    bounds, needs = [r[0] for r in rows], [r[1] for r in rows]
    (r0, r1), (n0, n1) = rows[rank]
    got = exchange_halo_rows(full[r0:r1].clone(), plan, rank, world, bounds=bounds, needs=needs)
    assert n0 <= r0 and n1 >= r1
    assert torch.equal(got, full[n0:n1]), f"rank {rank}: rows differ"
    # a dedup band reads less than a tile-row band: only the I - S rows above its first lattice row
    band = plan.dedup_band(world, rank)
    assert n0 == max(0, band.j0 * plan.stride - plan.off)
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_dedup_mode_input_rows_exchange_gloo(tmp_path, world):
    """Dedup mode's sharded loading: ranks own the raster rows they finalise (a partition of [0, H)) and fetch the rows
    their first lattice rows read from the rank above."""
    import torch.multiprocessing as mp
    from moonsuperresolution_b200.planner import Plan
    plan = Plan(2900, 37, 64, 16, 256, 4)
    own = [plan.dedup_rows(world, r)[0] for r in range(world)]
    assert own[0][0] == 0 and own[-1][1] == 2900 and all(a[1] == b[0] for a, b in zip(own[:-1], own[1:]))
    mp.spawn(_dedup_halo_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
