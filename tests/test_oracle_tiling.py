"""Pins oracle/tiling.py (numpy restatement) against outputs of the reference's own class (tests/golden)."""
import hashlib

import numpy as np
import pytest

import golden_inputs
import toy_models
from oracle import tiling


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("i", [32, 48, 64, 256, 512])
def test_blend_weight_table_bit_exact(golden, i):
    w = tiling.blend_weights(i)
    assert w.dtype == np.float64 and w.shape == (i - 2 * (i // 16),) * 2
    assert sha(w) == str(golden[f"weights/{i}/sha"])
    assert w.max() == 1.0000001 and abs(w.min() - golden[f"weights/{i}/minmax"][0]) == 0


def test_weight_hashes_match_survey_appendix_e(golden):
    assert str(golden["weights/256/sha"]).startswith("d8353b281465d155")
    assert str(golden["weights/512/sha"]).startswith("4f79ff011fb72781")


@pytest.mark.parametrize("name", list(golden_inputs.CASES))
def test_process_map_bit_exact_vs_reference(golden, name):
    case = golden_inputs.CASES[name]
    dem, img = golden_inputs.make_rasters(case)
    model = getattr(toy_models, case["model"])
    geo = tiling.Geometry(case["H"], case["W"], case["I"], case["S"], case["T"])
    assert (geo.canvas_h, geo.canvas_w) == tuple(golden[f"{name}/canvas"])
    assert (geo.pad_x, geo.pad_y) == tuple(golden[f"{name}/pad"])
    assert len(tiling.tile_list(geo)) == int(golden[f"{name}/ntiles"])
    mean, std, good = tiling.process_map(dem, img, case["I"], case["S"], case["B"], case["T"], case["NV"], model)
    assert mean.shape == std.shape == good.shape == (case["H"], case["W"])
    assert int(good.sum()) == int(golden[f"{name}/good_count"])
    assert sha(good) == str(golden[f"{name}/sha_good"])
    if case.get("store_full"):
        np.testing.assert_array_equal(mean, golden[f"{name}/mean"])
        np.testing.assert_array_equal(std, golden[f"{name}/std"])
    assert sha(mean) == str(golden[f"{name}/sha_mean"])
    assert sha(std) == str(golden[f"{name}/sha_std"])


def test_identity_round_trip_facts(golden):
    """identity model reproduces the DEM on good pixels; bbox and NV fill."""
    case = golden_inputs.CASES["identity_700x900"]
    dem, img = golden_inputs.make_rasters(case)
    mean, std, good = tiling.process_map(dem, img, 64, 8, 16, 256, case["NV"], None)
    g = good.astype(bool)
    rows, cols = np.where(g.any(1))[0], np.where(g.any(0))[0]
    assert (rows[0], rows[-1], cols[0], cols[-1]) == (4, 691, 4, 891)
    assert abs(g.mean() - 0.96975) < 1e-4
    assert np.abs(mean[g] - dem[g]).max() <= 1.3e-4
    assert 0 <= std[g].min() and std[g].max() <= 1.3e-4
    assert (mean[~g] == case["NV"]).all() and (std[~g] == case["NV"]).all()


def test_slot_counts_interior_tile():
    """SURVEY.md App. E interior-tile slot counts."""
    for (i, s, b), (patches, slots, batches) in {(512, 64, 12): (529, 540, 45), (512, 128, 16): (121, 128, 8),
                                                 (256, 32, 16): (1521, 1536, 96)}.items():
        geo = tiling.Geometry(3072, 3072, i, s, 1024)
        keys = list(tiling.patch_origins(geo, 1024, 1024))
        assert len(keys) == patches
        plan = tiling.batch_plan(keys, b)
        assert len(plan) == batches and sum(len(p) for p in plan) == slots


@pytest.mark.parametrize("name", ["wobble_200x260", "wobble_1100x1300"])
def test_dedup_oracle_equals_tile_by_tile_for_per_sample_models(name):
    """The dedup-mode oracle (global lattice, each patch generated once) is pinned by the pinned tile-by-tile oracle:
    with a model that does not depend on batch composition every pixel receives the same patches in the same order, so
    the rasters are bit-identical -- also when the lattice rows are cut into per-rank bands."""
    import toy_models
    from moonsuperresolution_b200.planner import Plan
    case = golden_inputs.CASES[name]
    dem, img = golden_inputs.make_rasters(case)
    args = (dem, img, case["I"], case["S"], case["B"], case["T"], case["NV"], toy_models.ripple)
    with np.errstate(all="ignore"):
        ref = tiling.process_map(*args)
        plan = Plan(case["H"], case["W"], case["I"], case["S"], case["T"], case["B"])
        assert tiling.lattice_counts(tiling.Geometry(case["H"], case["W"], case["I"], case["S"], case["T"])) == \
            plan.lattice_counts()
        for world in (1, 3):
            bands = [(plan.dedup_band(world, r).j0, plan.dedup_band(world, r).j1) for r in range(world)]
            got, batches = tiling.process_map_dedup(*args, row_bands=bands, return_plan=True)
            for a, b in zip(got, ref):
                np.testing.assert_array_equal(a, b)
            assert len(batches) == world


def test_repeats_one_is_the_reference_path_and_more_repeats_shrink_nothing():
    """oracle ``repeats``: 1 reproduces the pinned path bit for bit; with a deterministic model R identical generations
    leave the mean where it was (up to float32 rounding of the running update) and keep ``good`` unchanged."""
    import toy_models
    case = golden_inputs.CASES["wobble_200x260"]
    dem, img = golden_inputs.make_rasters(case)
    args = (dem, img, case["I"], case["S"], case["B"], case["T"], case["NV"], toy_models.ripple)
    with np.errstate(all="ignore"):
        base = tiling.process_map(*args)
        one = tiling.process_map(*args, repeats=1)
        three = tiling.process_map(*args, repeats=3)
        flick = tiling.process_map(*args[:-1], toy_models.Flicker(), repeats=3)
    for a, b in zip(base, one):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(three[2], base[2])
    g = base[2].astype(bool)
    assert np.abs(three[0][g] - base[0][g]).max() <= 1e-3 * np.abs(base[0][g]).max()
    assert flick[1][g].mean() > base[1][g].mean()          # repetitions that disagree add to the uncertainty
