"""Host-side integer planning vs the oracle restatement of the reference loops, plus the band sharding."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from moonsuperresolution_b200.distributed import assemble_bands, band_of_rank
from moonsuperresolution_b200.planner import PAD_SLOT, Plan, plan_batches, shard_tiles
from oracle import tiling as OT


@st.composite
def geometry(draw):
    i = draw(st.sampled_from([16, 32, 48, 64, 128, 256, 512]))
    t = draw(st.sampled_from([128, 256, 512, 1024]))
    divisors = [s for s in range(1, i + 1) if (t + i) % s == 0 and s >= 4]
    s = draw(st.sampled_from(divisors))
    h = draw(st.integers(1, 3000))
    w = draw(st.integers(1, 3000))
    b = draw(st.integers(1, 20))
    return h, w, i, s, t, b


@given(geometry())
@settings(max_examples=150, deadline=None)
def test_plan_matches_oracle_geometry(g):
    h, w, i, s, t, b = g
    plan = Plan(h, w, i, s, t, b)
    geo = OT.Geometry(h, w, i, s, t)
    assert (plan.canvas_h, plan.canvas_w, plan.pad_x, plan.pad_y, plan.off, plan.purge) == \
           (geo.canvas_h, geo.canvas_w, geo.pad_x, geo.pad_y, geo.off, geo.purge)
    assert plan.tiles() == OT.tile_list(geo)
    px, py = plan.tiles()[-1]
    assert [tuple(r) for r in plan.tile_patch_origins(px, py)] == list(OT.patch_origins(geo, px, py))
    assert plan.lattice_side ** 2 == len(list(OT.patch_origins(geo, px, py)))
    # every tile reads inside the canvas
    assert px + t + 2 * plan.off <= plan.canvas_w and py + t + 2 * plan.off <= plan.canvas_h


@given(st.integers(0, 200), st.integers(1, 17))
def test_batch_plan_matches_oracle(n, b):
    keys = [(k, 2 * k) for k in range(n)]
    assert plan_batches(keys, b) == OT.batch_plan(keys, b)
    plan = Plan(100, 100, 64, 8, 256, b)
    assert plan.batch_slots(n) == sum(len(x) for x in OT.batch_plan(keys, b))
    if n % b:
        assert plan_batches(keys, b)[-1][-1] == PAD_SLOT


def test_invalid_parameters_are_rejected():
    with pytest.raises(ValueError):
        Plan(100, 100, 8, 4, 256, 1)          # purge would be 0 (process_full_tiles.py:392-393)
    with pytest.raises(ValueError):
        Plan(100, 100, 256, 224, 1024, 1)     # last patch overruns the accumulator
    with pytest.raises(ValueError):
        Plan(100, 1000, 64, 8, 768, 1)        # tile beyond the 1024-quantised canvas
    with pytest.raises(ValueError):
        Plan(0, 10, 64, 8, 256, 1)


@given(st.integers(1, 80), st.integers(1, 9))
def test_shard_tiles_partitions_contiguously(n, world):
    tiles = [(0, k) for k in range(n)]
    owned = [shard_tiles(tiles, world, r) for r in range(world)]
    flat = [t for o in owned for t in o]
    assert flat == tiles                                   # disjoint, complete, order preserving, contiguous
    sizes = [len(o) for o in owned]
    assert max(sizes) - min(sizes) <= 1


def test_shard_tiles_balances_cost():
    tiles = [(0, k) for k in range(10)]
    cost = [1] * 5 + [9] * 5
    owned = [shard_tiles(tiles, 2, r, cost) for r in range(2)]
    assert owned[0] + owned[1] == tiles
    loads = [sum(cost[t[1]] for t in o) for o in owned]
    assert abs(loads[0] - loads[1]) <= 9


def test_bands_cover_the_raster_once():
    plan = Plan(15000, 2000, 512, 128, 1024, 16)
    for world in (1, 2, 4, 8):
        covered = np.zeros(plan.height, np.int32)
        seen = []
        for r in range(world):
            tiles, r0, r1 = band_of_rank(plan, world, r)
            covered[r0:r1] += 1
            seen += tiles
        assert (covered == 1).all()
        assert sorted(seen) == sorted(plan.tiles())
    parts = [(0, np.ones((3, 4), np.float32)), (3, 2 * np.ones((2, 4), np.float32))]
    out = assemble_bands(parts, 5, 4, np.float32)
    assert out[:3].min() == 1 and out[3:].min() == 2


def test_parse_args_mirrors_the_reference_cli():
    """Flag names and defaults of process_full_tiles.py:68-127; unknown flags are ignored (:114)."""
    from moonsuperresolution_b200 import DSRConfig, parse_args
    cfg = parse_args(["--source_folder_path", "in", "--map_name", "m", "--save_path", "out", "--whatever", "1"])
    ref = DSRConfig()
    for f in ("image_size", "stride", "batch_size", "tile_size", "no_value", "upsample_factor", "ortho_image_name",
              "dem_name", "model_path"):
        assert getattr(cfg, f) == getattr(ref, f)
    assert (cfg.image_size, cfg.stride, cfg.batch_size, cfg.tile_size, cfg.no_value) == (256, 32, 16, 1024, -32768.0)
    cfg = parse_args(["--source_folder_path", "in", "--map_name", "m", "--save_path", "out", "--image_size", "512",
                      "--stride", "64", "--batch_size", "12", "--model_path", "w/"])
    assert (cfg.image_size, cfg.stride, cfg.batch_size, cfg.model_path) == (512, 64, 12, "w/")
    with pytest.raises(SystemExit):
        parse_args(["--map_name", "m"])                     # required flags, as in the reference


@pytest.mark.parametrize("dims,world", [((15000, 70000, 512, 128, 1024, 16), 8), ((8192, 8192, 512, 128, 1024, 16), 2),
                                        ((200, 260, 32, 8, 128, 5), 3), ((4096, 4096, 256, 32, 1024, 16), 4),
                                        ((200, 200, 32, 32, 128, 3), 2)])
def test_dedup_bands_partition_lattice_and_canvas(dims, world):
    """Dedup mode (mode B): lattice rows and finalised canvas rows are partitioned once, every band reads a
    superset of what it finalises, and the strip a rank sends is the strip the next one expects."""
    plan = Plan(*dims)
    gy, gx = plan.lattice_counts()
    i, s, p = plan.image_size, plan.stride, plan.purge
    assert (gy - 1) * s + i <= plan.canvas_h and (gx - 1) * s + i <= plan.canvas_w
    bands = [plan.dedup_band(world, r) for r in range(world)]
    assert bands[0].j0 == 0 and bands[-1].j1 == gy and bands[0].out[0] == 0 and bands[-1].out[1] == plan.canvas_h
    rows_out = 0
    for a, b in zip(bands[:-1], bands[1:]):
        assert a.j1 == b.j0 and a.out[1] == b.out[0]
        assert a.seam_out == b.seam_in
        if a.seam_out is not None:
            # the strip lies inside both ranks' accumulators and below what the sender finalises
            assert a.read[0] <= a.seam_out[0] and a.seam_out[1] <= a.read[1]
            assert b.read[0] <= b.seam_in[0] and b.seam_in[1] <= b.read[1]
            assert a.seam_out[0] == a.out[1]
            # it ends where the sender's last patch row stops contributing
            assert a.seam_out[1] == (a.j1 - 1) * s + i - p
    for k, band in enumerate(bands):
        assert band.j1 - band.j0 >= 1
        lo, hi = band.out
        if k > 0:
            assert band.read[0] <= lo
        if k < world - 1:
            assert hi <= band.read[1]
        r0, r1 = band.raster_rows(band.out, plan.off, plan.height)
        rows_out += r1 - r0
    assert rows_out == plan.height
    # work balance: lattice rows whose patches lie inside the raster are spread evenly
    inside = lambda j: j * s >= plan.off and j * s + i <= plan.off + plan.height
    counts = [sum(inside(j) for j in range(b.j0, b.j1)) for b in bands]
    assert max(counts) - min(counts) <= 1


def test_dedup_rejects_bad_geometry():
    with pytest.raises(ValueError):
        Plan(150, 330, 24, 16, 120, 7).lattice_counts()              # S divides T + I but not T
    with pytest.raises(ValueError):
        Plan(200, 260, 32, 8, 128, 5).dedup_band(16, 0)              # bands thinner than a patch
